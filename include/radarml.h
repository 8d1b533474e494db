/*
 * radarml.h — C ABI of libradarml.so: the B200-native replacement for the per-scan
 * classification hot path of goruck/radar-ml (reference @ 6f7af2c).
 *
 * The reference is pure in-process Python and has no FFI of its own; each
 * entry point below cites the reference call site whose arithmetic it replaces.  The Python
 * host (radar_ml_b200/) binds these with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add to predict.py.
 *
 * Conventions
 *   - every function returns 0 (RML_OK) or a negative rml_status; rml_last_error() gives text
 *   - pointers named *_dev are device pointers on the context's GPU, *_host are host pointers
 *   - the caller owns every buffer; the library owns only the context and its model copy
 *   - device entry points are asynchronous on the given cudaStream_t and never synchronise
 *   - a context is bound to one device and is not thread-safe; use one context per GPU
 *   - there is NO CPU fallback anywhere: without a GPU, rml_create fails with RML_E_CUDA
 */
#ifndef RADARML_H_
#define RADARML_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rml_ctx rml_ctx;
typedef void* rml_stream; /* cudaStream_t */

typedef enum rml_status {
  RML_OK = 0,
  RML_E_INVALID = -1,      /* bad argument / shape */
  RML_E_CUDA = -2,         /* CUDA runtime or driver error, no usable device */
  RML_E_UNSUPPORTED = -3,  /* e.g. zoom != 1.0 projections requested of the fused kernel */
  RML_E_NOMODEL = -4,      /* scoring requested before rml_load_* */
  RML_E_NONINTEGRAL = -5,  /* u8 fast path saw a value that is not an integer in [0,255] */
  RML_E_RANGE = -6         /* digit path saw a feature outside [0, 256/feature_scale) */
} rml_status;

/* projection mode: predict.py:102-107 takes SLICES through the target voxel (i,j,k);
 * BASELINE.json north_star reduces with an axis MAX.  Output shapes are identical. */
enum { RML_MODE_MAX = 0, RML_MODE_SLICE = 1 };

/* proj_mask bits, in the reference's tuple order (xz, yz, xy): common.py:40, predict.py:113 */
enum { RML_MASK_XZ = 1, RML_MASK_YZ = 2, RML_MASK_XY = 4, RML_MASK_ALL = 7 };

/* feature dtypes produced by rml_project / consumed by rml_score */
enum {
  RML_F32 = 0, /* float32 (n,F) exactly as common.process_samples returns (common.py:149) */
  RML_U8 = 1,  /* raw integer sensor value 0..255, K-padded rows, operand layout of the
                  tensor-core scorer; the /255 scale is folded into the scorer */
  RML_F32_EXACT = 2 /* rml_score only: float32 features through the float64 CUDA-core scorer (any
                  value range); plain RML_F32 uses the exact multi-digit tensor-core scorer, which
                  needs every feature and support-vector component in [0, 256/feature_scale) */
};

/* ---- lifetime ------------------------------------------------------------------------- */
int rml_create(int device, rml_ctx** out);
int rml_destroy(rml_ctx* ctx);
const char* rml_last_error(const rml_ctx* ctx); /* ctx may be NULL: last create() error */
int rml_version(void);

/* ---- layout: common.py:25-27 arena -> cube dims (size_x,size_y,size_z) = (22,31,176) ---- */
int rml_set_arena(rml_ctx* ctx, int size_x, int size_y, int size_z);
/* spherical bounds used by calculate_matrix_indices (common.py:25-27 R/THETA/PHI MIN,MAX) */
int rml_set_arena_bounds(rml_ctx* ctx, double r_min, double r_max, double theta_min,
                         double theta_max, double phi_min, double phi_max);
int rml_feature_len(const rml_ctx* ctx, uint32_t mask);    /* F for a mask (10010 for all)   */
int rml_feature_stride(const rml_ctx* ctx, uint32_t mask, int dtype); /* elements per row    */

/* Affine (x - offset) / scale applied by rml_project for RML_F32 output, as an IEEE float32 true
 * division.  The scalar form covers common.py:148 (/255., offset 0) and dnn.py:202-205
 * ((p-127.5)/127.5). */
int rml_set_affine(rml_ctx* ctx, float offset, float scale, int enabled);
/* Per-feature form, SURVEY.md §8b: offset_host[F], scale_host[F] (a fitted StandardScaler's mean_
 * and scale_ — what BASELINE.json north_star calls "StandardScaler-normalised"; the reference
 * itself only divides by RADAR_MAX, common.py:148, which stays the default).  NULL, NULL restores
 * that default.  With tables loaded every float32 feature producer (rml_project, rml_predict*,
 * rml_process_samples*) emits (x - offset[f]) / scale[f]; the integer u8 operand rows cannot
 * carry a per-feature scale, so rml_predict* switch to float32 features and the exact
 * multi-digit tensor-core scorer (or the float64 scorer) — load the affine BEFORE the model. */
int rml_load_affine(rml_ctx* ctx, const float* offset_host, const float* scale_host, int F);
/* force_f32 = 1: rml_predict* emit float32 features and score them with the multi-digit /
 * float64 scorer even when the model qualifies for the u8 path — the path real-valued cubes
 * need (the u8 path reports them with RML_E_NONINTEGRAL instead of scoring them). 0 = automatic. */
int rml_set_precision(rml_ctx* ctx, int force_f32);

/* ---- model load: what predict.py:224-225 unpickles ------------------------------------ */
/* CalibratedClassifierCV(prefit SVC rbf) built at train.py:478-479, 723-724.
 * sv (n_sv,F) row-major, dual_coef (n_classes-1,n_sv), rho (n_classes*(n_classes-1)/2)
 * (= -intercept_), n_support (n_classes), platt a/b per class in estimator.classes_ order
 * (one pair when n_classes == 2).  All host pointers, float64 like sklearn holds them.
 * feature_scale: the value the training features were divided by (255, common.py:31). */
int rml_load_svc_rbf(rml_ctx* ctx, int n_classes, int F, int n_sv, const int32_t* n_support,
                     const double* sv_host, const double* dual_coef_host, const double* rho_host,
                     double gamma, const double* platt_a_host, const double* platt_b_host,
                     double feature_scale);
/* SGDClassifier(loss='log') alternative, train.py:368-369 (deployed in predict.log:9). */
int rml_load_linear(rml_ctx* ctx, int n_classes, int F, const double* coef_host,
                    const double* intercept_host, const double* platt_a_host,
                    const double* platt_b_host, double feature_scale);
/* 1 when the loaded SVC's support vectors are integers/feature_scale (u8 tensor path usable) */
int rml_model_is_integral(const rml_ctx* ctx);

/* ---- K1: predict.py:102-107 + common.py:141-149 (projection, concat, scale) ------------ */
/* cubes_dev: float32 [B][size_x][size_y][size_z] (predict.py:91 layout, z fastest).
 * ijk_dev: int32 [B][3] target voxel for RML_MODE_SLICE (common.py:106-121), else NULL.
 * feats_dev: [B][rml_feature_stride()] of dtype.  norms_dev: int32 [B] sum of squares of the
 * u8 row (RML_U8 only, may be NULL). */
int rml_project(rml_ctx* ctx, const float* cubes_dev, int64_t B, int mode, const int32_t* ijk_dev,
                uint32_t mask, int dtype, void* feats_dev, int32_t* norms_dev, rml_stream stream);
/* common.process_samples on already-extracted projections (common.py:123-149, zoom 1.0):
 * xz [B][sx][sz], yz [B][sy][sz], xy [B][sx][sy] float32 device arrays (NULL if masked out).
 * scale != 0 applies the rml_set_affine transform (default (x-0)/255 = common.py:148). */
int rml_process_samples(rml_ctx* ctx, const float* xz_dev, const float* yz_dev,
                        const float* xy_dev, int64_t B, uint32_t mask, int scale,
                        float* feats_dev, rml_stream stream);
/* common.calculate_matrix_indices (common.py:106-121) for B targets: xyz_dev float64 [B][3] */
int rml_matrix_indices(rml_ctx* ctx, const double* xyz_dev, int64_t B, int32_t* ijk_dev,
                       rml_stream stream);

/* common.py:45-80 DerivedTarget.get_derived_targets: per scan, the num_targets indices with the
 * largest axis sums of the raw cube along theta (i), phi (j) and r (k), ascending by sum like
 * np.argsort (the last triple is the strongest).  ijk_dev int32 [B][num_targets][3];
 * sums_dev (nullable) float32 [B][sx+sy+sz] = theta | phi | r sums. */
int rml_derive_targets(rml_ctx* ctx, const float* cubes_dev, int64_t B, int num_targets,
                       int32_t* ijk_dev, float* sums_dev, rml_stream stream);
/* rml_project(RML_MODE_MAX) and rml_derive_targets in ONE pass over the cubes ("the
 * axis sums can share K1's single HBM pass"): the projection kernel's row warps accumulate the theta /
 * phi / r sums from the voxels they already hold in registers, its flusher warp ranks them.  Same
 * outputs as the two separate calls (identical for integer-valued cubes; for real-valued cubes the
 * float32 sums differ in the last bits because the summation order differs).  Default arena; other
 * arenas run the two kernels back to back. */
int rml_project_derive(rml_ctx* ctx, const float* cubes_dev, int64_t B, uint32_t mask, int dtype,
                       void* feats_dev, int32_t* norms_dev, int num_targets, int32_t* ijk_dev,
                       float* sums_dev, rml_stream stream);
/* common.py:143-144 scipy.ndimage.zoom(p, proj_zoom[i]) (order-3 spline) for arenas that differ
 * from the training arena (predict.py:34-54, README.md:207).  For fixed sizes the zoom is a
 * separable linear operator: a_rows [out_h][in_h], a_cols [out_w][in_w] (HOST float64,
 * extracted from scipy by the caller).  proj: 0 xz, 1 yz, 2 xy. */
int rml_set_zoom(rml_ctx* ctx, int proj, int in_h, int in_w, int out_h, int out_w,
                 const double* a_rows_host, const double* a_cols_host);
int rml_zoom_feature_len(const rml_ctx* ctx, uint32_t mask);
/* common.process_samples with zoom: projection p of scan b starts at p_dev + b*stride_p
 * (elements), so both separate [B][h][w] arrays and rml_project feature rows can be fed.
 * feats_dev float32 [B][rml_zoom_feature_len(mask)], scaled by the rml_set_affine transform
 * when scale != 0. */
int rml_process_samples_zoom(rml_ctx* ctx, const float* xz_dev, int64_t stride_xz,
                             const float* yz_dev, int64_t stride_yz, const float* xy_dev,
                             int64_t stride_xy, int64_t B, uint32_t mask, int scale,
                             float* feats_dev, rml_stream stream);

/* ---- K2: predict.py:56-70 classifier() -> model.predict_proba chain -------------------- */
/* feats_dev as written by rml_project (dtype RML_U8 needs norms_dev) or float32 (n,F)
 * features already scaled like common.process_samples(scale=True).
 * proba_dev float32 [B][n_classes]; decision_dev float32 [B][n_classes] OvR decision values
 * (nullable; [B] when n_classes == 2); label_dev int32 [B] = argmax; known_dev uint8 [B] =
 * (max proba >= min_proba), i.e. name != 'Unknown' at predict.py:65-68. */
int rml_score(rml_ctx* ctx, const void* feats_dev, int dtype, const int32_t* norms_dev, int64_t B,
              double min_proba, float* proba_dev, float* decision_dev, int32_t* label_dev,
              uint8_t* known_dev, rml_stream stream);

/* (n,F) float32 features scaled like common.process_samples(scale=True) -> the u8 operand
 * rows + norms rml_score(RML_U8) consumes.  Values that are not exactly
 * float32(u)/float32(feature_scale), u integer in [0,255], are reported by rml_check_status
 * (RML_E_NONINTEGRAL): score those features as RML_F32 instead. */
int rml_quantize_features(rml_ctx* ctx, const float* feats_dev, int64_t B, int F,
                          uint8_t* feats_u8_dev, int32_t* norms_dev, rml_stream stream);

/* ---- K1 -> K2 fused pipeline: one predict.py:93-119 iteration for B scans -------------- */
/* workspace: rml_predict_workspace_bytes(ctx, B) bytes of 256-byte aligned device memory for the
 * loaded model (feature staging, norms, digit planes of the float32 path, tile counters of the
 * co-resident pipeline): rml_predict* allocate nothing.  label_dev may point into a larger
 * gather buffer (see rml_allgather_labels). */
size_t rml_predict_workspace_bytes(const rml_ctx* ctx, int64_t B);
int rml_predict(rml_ctx* ctx, const float* cubes_dev, int64_t B, int mode, const int32_t* ijk_dev,
                uint32_t mask, double min_proba, void* workspace_dev, float* proba_dev,
                int32_t* label_dev, uint8_t* known_dev, rml_stream stream);
/* Same, HOST buffers in and out (cubes_host pinned or pageable float32; results to host).
 * Streams the batch through the GPU in chunks with H2D / compute / D2H overlapped and
 * returns after the results are in host memory.  This is the e2e entry bench.py times. */
int rml_predict_host(rml_ctx* ctx, const float* cubes_host, int64_t B, int mode,
                     const int32_t* ijk_host, uint32_t mask, double min_proba, float* proba_host,
                     int32_t* label_host, uint8_t* known_host);

/* rml_predict_host and integer-valued float32 cubes (predict.py:90-91 widens the sensor's integers
 * 0..255 to float32): for a model on the integer path each chunk is converted back to bytes on the
 * host (a pool of threads, every value checked to be exactly an integer in [0,255]) and a quarter
 * of the bytes crosses PCIe; the device runs the uint8-cube kernels, whose results are identical bit
 * for bit.  A chunk holding any other value is copied as float32 as before (two such chunks switch
 * the check off).  ON by default (env RML_HOST_NARROW=0 or enabled = 0 disables): the conversion of
 * chunk n+1 overlaps the copy and the kernels of chunk n — results leave through pinned mirrors, so
 * no copy blocks the calling thread — and it pays where the host converts faster than the bus moves
 * float32 bytes: on the B200 box (16 vCPUs, PCIe 5 x16) 65-95 GB/s of float32 input against 53 GB/s
 * over the bus, 113 k -> 140-145 k scans/s end to end (bench.py e2e / e2e.plain_copy).  It switches
 * itself off when three consecutive chunks convert slower than min_gbs (default 57).  threads = 0:
 * one per CPU of the process's affinity mask.  Call before rml_reserve(RML_RESERVE_HOST). */
int rml_set_host_narrowing(rml_ctx* ctx, int enabled, int threads, double min_gbs);
/* the conversion itself (host only, no GPU needed): dst_host[i] = (uint8_t)src_host[i]; RML_OK when every
 * value is exactly an integer in [0,255], RML_E_NONINTEGRAL otherwise.  dst_host 32-byte aligned. */
int rml_host_narrow_f32_to_u8(const float* src_host, uint8_t* dst_host, int64_t n, int threads);
/* what the last rml_predict_host moved: bytes copied host->device, scans that crossed as bytes,
 * float32 input rate of the last conversion (GB/s), pool threads, 1 while narrowing is active */
int rml_last_host_transfer(const rml_ctx* ctx, int64_t* h2d_bytes, int64_t* narrowed_scans,
                           double* convert_gbs, int* threads, int* active);

/* Buffers the library owns are created here, never inside a hot entry point (SURVEY.md §8b
 * ownership row): call after the model is loaded, again after loading another model.
 *   RML_RESERVE_SCORE    digit-plane scratch so rml_score accepts float32 rows of up to max_batch
 *   RML_RESERVE_HOST     chunked staging of rml_predict_host (float32 cubes)
 *   RML_RESERVE_HOST_U8  the same for rml_predict_host_u8
 *   RML_RESERVE_SMALL    staging of rml_score_host / rml_predict_targets_host for max_batch rows */
enum { RML_RESERVE_SCORE = 1, RML_RESERVE_HOST = 2, RML_RESERVE_HOST_U8 = 4, RML_RESERVE_SMALL = 8 };
int rml_reserve(rml_ctx* ctx, int64_t max_batch, int flags);

/* predict.py:56-70 classifier() for HOST features: feats_host float32 [B][F] as
 * common.process_samples(scale=True) returns them -> predict_proba, argmax, >= min_proba.  The
 * scorer is chosen on the device: u8 tensor-core scorer for integral rows of an integral model,
 * else the multi-digit tensor-core scorer, else float64; one synchronisation in the common case. */
int rml_score_host(rml_ctx* ctx, const float* feats_host, int64_t B, double min_proba,
                   float* proba_host, int32_t* label_host, uint8_t* known_host);
/* One predict.py:93-119 iteration — the reference's live loop body: ONE raw cube (predict.py:90-91,
 * float32 [size_x][size_y][size_z], host) and the T targets GetSensorTargets reported for it
 * (ijk_host int32 [T][3], common.py:106-121).  The cube is uploaded once. */
int rml_predict_targets_host(rml_ctx* ctx, const float* cube_host, int T, const int32_t* ijk_host,
                             uint32_t mask, double min_proba, float* proba_host,
                             int32_t* label_host, uint8_t* known_host);

/* ---- multi-GPU: the one exchange of the path ----------------------------- */
/* One context per GPU/process.  rank 0 calls rml_comm_unique_id (128 bytes), hands the id to the
 * other ranks by any host channel, every rank calls rml_comm_init.  NCCL is dlopen'ed
 * (libnccl.so.2 — the copy already in the process if there is one, e.g. torch's; env
 * RML_NCCL_LIB overrides); a single-GPU process never needs it. */
int rml_comm_unique_id(rml_ctx* ctx, void* id128_host);
int rml_comm_init(rml_ctx* ctx, int rank, int world, const void* id128_host);
int rml_comm_destroy(rml_ctx* ctx);
/* ncclAllGather of int32 labels on `stream`: recv_dev [world][count].  send_dev may be
 * recv_dev + rank*count (in place): pass that slice as rml_predict's label_dev and the scorer
 * writes straight into the gather buffer.  world == 1 without a communicator degenerates to a copy. */
int rml_allgather_labels(rml_ctx* ctx, const int32_t* send_dev, int32_t* recv_dev, int64_t count,
                         rml_stream stream);

/* ---- uint8 cubes: the sensor's integers kept as bytes ----------------------------------- */
/* predict.py:90-91 widens the Walabot's integer voxels (0..255, ground_truth_samples.py:352)
 * with np.array(raw_image, dtype=np.float32).  A caller that holds them as uint8
 * (np.array(raw_image, dtype=np.uint8)) moves a quarter of the bytes over PCIe and HBM; the
 * results are those of the float32 entry points on the widened cube, bit for bit.
 * cubes: uint8 [B][size_x][size_y][size_z], same axis order; everything else as in
 * rml_project / rml_predict / rml_predict_host / rml_net_predict. */
int rml_project_u8(rml_ctx* ctx, const uint8_t* cubes_dev, int64_t B, int mode,
                   const int32_t* ijk_dev, uint32_t mask, int dtype, void* feats_dev,
                   int32_t* norms_dev, rml_stream stream);
int rml_predict_u8(rml_ctx* ctx, const uint8_t* cubes_dev, int64_t B, int mode,
                   const int32_t* ijk_dev, uint32_t mask, double min_proba, void* workspace_dev,
                   float* proba_dev, int32_t* label_dev, uint8_t* known_dev, rml_stream stream);
int rml_predict_host_u8(rml_ctx* ctx, const uint8_t* cubes_host, int64_t B, int mode,
                        const int32_t* ijk_host, uint32_t mask, double min_proba,
                        float* proba_host, int32_t* label_host, uint8_t* known_host);

/* ---- dnn.py / sgan.py classifier forward ------------------------ */
/* Load sequence: begin -> resize tables x3 -> conv layers (per layer, per branch) -> dense ->
 * finish.  All pointers are HOST pointers.  BatchNorm (sgan.py:138,144,150,190,195) is folded
 * into the preceding kernel/bias by the caller (inference form, epsilon 1e-3); Dropout is the
 * identity at inference.
 *   resize_to   80 (dnn.py:33) or 128 (sgan.py:39);  head 0 = softmax (dnn.py:85, sgan c_model
 *   sgan.py:205), 1 = Z/(Z+1), Z = sum exp(logit) (sgan d_model, sgan.py:125-129, 210-213)
 *   act: 0 none, 1 ReLU (dnn.py:48-52), 2 LeakyReLU(alpha) (sgan.py:141) */
int rml_net_begin(rml_ctx* ctx, int resize_to, int n_classes, int head, float lrelu_alpha);
/* Pillow Resample.c BICUBIC coefficient tables of branch b (0 xz, 1 yz, 2 xy): kh [R][ksh]
 * with bounds bh [R][2] = (first input column, taps) for the horizontal pass, kv/bv vertical */
int rml_net_set_resize_tables(rml_ctx* ctx, int branch, int ksh, const double* kh_host,
                              const int32_t* bh_host, int ksv, const double* kv_host,
                              const int32_t* bv_host);
/* Conv2D(cout, 3x3, strides 2, 'same') of tower `branch`, layer index from 0: kernel in Keras
 * (kh,kw,cin,cout) order, dnn.py:45-52 / sgan.py:132-154 */
int rml_net_add_conv(rml_ctx* ctx, int layer, int branch, int cin, int cout, int act,
                     const float* w_hwio_host, const float* bias_host);
/* Dense 64 -> Dense 64 -> Dense C (dnn.py:80-85, sgan.py:188-202).  w1t: bf16 bits [64][K],
 * K ordered [branch][h][w][c] (the caller permutes Keras' Flatten order (h,w,96)); w2 [64][64]
 * and w3 [64][C] input-major float32. */
int rml_net_set_dense(rml_ctx* ctx, int K, const uint16_t* w1t_bf16_host, const float* b1_host,
                      int act1, const float* w2_host, const float* b2_host, int act2,
                      const float* w3_host, const float* b3_host);
int rml_net_finish(rml_ctx* ctx);
/* 1 when the conv layers after the first run as tcgen05 implicit GEMMs on bf16 activations
 * (every such layer has Cin % 64 == 0, Cout % 16 == 0, Cout <= 128); env RML_IGEMM=0 forces the
 * fp32 CUDA-core towers. */
int rml_net_uses_igemm(const rml_ctx* ctx);
size_t rml_net_workspace_bytes(const rml_ctx* ctx, int64_t chunk_scans);
/* feats_dev: float32 [B][10010] projections scaled (p-127.5)/127.5 (rml_project with
 * rml_set_affine(127.5, 127.5, 1)).  The batch is processed in chunks that fit the workspace.
 * proba_dev [B][C], logits_dev [B][C] nullable, label_dev [B]. */
int rml_net_forward(rml_ctx* ctx, const float* feats_dev, int64_t B, void* workspace_dev,
                    size_t workspace_bytes, float* proba_dev, float* logits_dev, int32_t* label_dev,
                    rml_stream stream);
/* dnn.py:240-254 alone: scaled projections -> images_dev float32 [B][3][R][R] */
int rml_net_resize(rml_ctx* ctx, const float* feats_dev, int64_t B, float* images_dev,
                   rml_stream stream);
/* Keras model.predict([XZ, YZ, XY]) on preprocessed inputs images_dev [B][3][R][R].
 * tower_bf16_dev (nullable): receives the flattened conv-tower output, bf16 bits [B][K] in
 * [branch][h][w][c] order — the operand the tensor-core dense stack consumed. */
int rml_net_forward_images(rml_ctx* ctx, const float* images_dev, int64_t B, void* workspace_dev,
                           size_t workspace_bytes, float* proba_dev, float* logits_dev,
                           int32_t* label_dev, uint16_t* tower_bf16_dev, rml_stream stream);
/* cubes -> projections -> scale -> resize -> towers -> dense stack (configs[2], configs[4]) */
int rml_net_predict(rml_ctx* ctx, const float* cubes_dev, int64_t B, int mode,
                    const int32_t* ijk_dev, void* workspace_dev, size_t workspace_bytes,
                    float* proba_dev, int32_t* label_dev, rml_stream stream);

int rml_net_predict_u8(rml_ctx* ctx, const uint8_t* cubes_dev, int64_t B, int mode,
                       const int32_t* ijk_dev, void* workspace_dev, size_t workspace_bytes,
                       float* proba_dev, int32_t* label_dev, rml_stream stream);

/* ---- status of the last asynchronous work (non-integral count seen by the u8 path) ----- */
int rml_check_status(rml_ctx* ctx, rml_stream stream); /* synchronises the stream */

/* rml_predict runs K1 and K2 as ONE device-side pipeline for large batches: the projection
 * kernel streams cubes on (SMs - k2_sms) SMs and counts finished scans per 128-scan tile; the
 * scorer is co-resident on the other k2_sms SMs and starts a tile the moment it is complete.
 * enabled=0 forces the serial K1 -> K2 order (also: env RML_FUSED=0, RML_K2_SMS, RML_FUSED_MIN_B). */
int rml_set_fused(rml_ctx* ctx, int enabled, int k2_sms, int64_t min_batch);
/* uint8 cubes: k2_sms > 0 selects the co-resident pipeline with that many scorer SMs, 0 (the
 * default) the serial K1 -> K2 order — the byte-SIMD projection kernel is issue-bound and wants
 * every SM, so serial measured faster (profiles/r1c_u8_sweep.txt).  Also env RML_K2_SMS_U8. */
int rml_set_fused_u8(rml_ctx* ctx, int k2_sms);
/* CUDA-event timing of the kernels inside the last rml_predict on this context: k1_ms =
 * projection kernel alone, total_ms = projection start -> scorer end.  Synchronises. */
int rml_enable_timing(rml_ctx* ctx, int enabled);
int rml_last_timing(rml_ctx* ctx, float* k1_ms, float* total_ms, int* fused);

/* number of kernels this library launched since create (bench.py "gpu_launches") */
int64_t rml_launch_count(const rml_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RADARML_H_ */
