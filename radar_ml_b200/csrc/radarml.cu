// libradarml.so — C ABI (include/radarml.h) over the sm_100a kernels.
// Host side: context + model residency, tensor-map encoding, launches, host-buffer pipeline.
#include "../../include/radarml.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <initializer_list>
#include <string>
#include <vector>

#include <dlfcn.h>

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include "k1_project.cuh"
#include "k1_project_u8.cuh"
#include "k2_score.cuh"
#include "k2_digits.cuh"
#include "k3_net.cuh"
#include "k6_tower.cuh"
#include "k0_extras.cuh"
#include "host_narrow.h"

using namespace rml;

namespace {

thread_local std::string g_create_error;

struct Model {
  int kind = 0;  // 0 none, 1 svc_rbf, 2 linear
  int C = 0, F = 0, n_sv = 0;
  bool integral = false;
  double gamma = 0, feature_scale = 255.0;
  int class_end[kMaxClasses] = {0};
  // device
  uint8_t* sv_u8 = nullptr;   // [n_sv_pad][kpad]
  int32_t* svnorm = nullptr;  // [n_sv]
  double* sv_f64 = nullptr;   // [n_sv][F]
  double* coef = nullptr;     // svc: [C-1][n_sv]; linear: [R][F]
  double* pairw = nullptr;    // svc: [C(C-1)/2][n_pad] per-pair SV weights for the tensor-core scorer
  double* rho = nullptr;      // svc: [NP]; linear: intercept [R]
  double* platt_a = nullptr;
  double* platt_b = nullptr;
  int kpad = 0, n_tile = 0, n_chunks = 0;
  CUtensorMap map_sv;
  // multi-digit exact path for non-integral values in [0, 256/scale)
  bool digits_ok = false;
  uint8_t* sv_digits = nullptr;   // [3][dg_pad][kpad]
  long long* svnorm64 = nullptr;  // [dg_pad]
  double* pairw_dg = nullptr;     // [NP][dg_pad]
  int dg_pad = 0, dg_chunks = 0;
  CUtensorMap map_sv_dg[3];
  double dg_scale = 255.0;        // fixed-point scale of the digit planes: X = round((x + shift_f) * dg_scale * 2^16)
  double* dg_shift = nullptr;     // device [F] shift_f (null: 0) — per-feature affine loaded before the model
  int i8_stages = 0;              // ring depth of k2_rbf_i8 for this model
};

constexpr int kHostBufs = 3;
struct HostPipe {
  int64_t chunk = 0;
  size_t cube_bytes = 0;   // bytes per cube the buffers were sized for
  void* cubes[kHostBufs] = {nullptr};
  int32_t* ijk[kHostBufs] = {nullptr};
  void* work[kHostBufs] = {nullptr};
  float* proba[kHostBufs] = {nullptr};
  int32_t* label[kHostBufs] = {nullptr};
  uint8_t* known[kHostBufs] = {nullptr};
  cudaStream_t stream[kHostBufs] = {nullptr};
  size_t work_bytes = 0;
  // results leave through pinned mirrors: a cudaMemcpyAsync into the caller's (pageable) arrays blocks the
  // host thread until the stream has drained, which serialised the host-side work of chunk n+1 (the
  // narrowing below) behind the copy and the kernels of chunk n
  unsigned char* res_pin[kHostBufs] = {nullptr};   // [chunk][C floats | int32 label | u8 known]
  cudaEvent_t res_ev[kHostBufs] = {nullptr};       // the D2H copies into res_pin[i] have finished
  int64_t res_at[kHostBufs] = {0}, res_n[kHostBufs] = {0};   // scans [res_at, res_at + res_n) wait in res_pin[i]
  int narrow_bad = 0;              // chunks that turned out not to be integral (two of them switch narrowing off)
  // host-side narrowing of integral float32 cubes to bytes before the H2D copy (host_narrow.h)
  uint8_t* narrow_pin[kHostBufs] = {nullptr};      // pinned staging, chunk x cube elements each
  cudaEvent_t narrow_ev[kHostBufs] = {nullptr};    // the H2D copy out of a staging buffer has finished
  rml_host::NarrowPool* pool = nullptr;
  int narrow_slow = 0;             // consecutive chunks whose conversion was slower than the float32 copy would be
  bool narrow_off = false;         // switched off by that measurement (until the next rml_reserve)
  double narrow_gbs = 0;           // float32 input rate of the last conversion
  int64_t last_h2d_bytes = 0, last_narrowed = 0;   // of the last rml_predict_host call
};

// scratch of the small host-buffer calls (rml_score_host / rml_predict_targets_host): one scan's
// cube, T <= kSmallMax feature rows and their results, plus pinned mirrors of the results
constexpr int64_t kSmallMax = 1024;
struct SmallPipe {
  int64_t cap = 0;            // rows
  int F = 0;
  void* cube = nullptr;       // one cube (float32)
  int32_t* ijk = nullptr;
  float* feats = nullptr;     // [cap][F] f32
  void* work = nullptr;       // predict workspace for cap rows
  size_t work_bytes = 0;
  float* proba = nullptr;
  int32_t* label = nullptr;
  uint8_t* known = nullptr;
  unsigned char* pinned = nullptr;   // host: proba | label | known | status(16)
  size_t pinned_bytes = 0;
};

// NCCL through dlopen (the torch-bundled libnccl.so.2 or the system one): the library must not
// link against a particular NCCL build, and a process that never calls rml_comm_init needs none.
typedef struct { char internal[128]; } rml_nccl_uid;
struct Nccl {
  void* handle = nullptr;
  int (*GetUniqueId)(rml_nccl_uid*) = nullptr;
  int (*CommInitRank)(void**, int, rml_nccl_uid, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;


struct NetConv {
  int cin = 0, cout = 0, act = 0;
  float* w[3] = {nullptr, nullptr, nullptr};
  float* bias[3] = {nullptr, nullptr, nullptr};
  std::vector<float> w_host[3], b_host[3];   // kept for the implicit-GEMM repack at finish
  uint16_t* wt_bf16 = nullptr;               // [3*cout][9*cin] bf16, K = tap*cin + ci
  float* bias3 = nullptr;                    // [3][cout]
  CUtensorMap map_w;
};
struct Net {
  bool ready = false;
  bool use_igemm = false;   // layers >= 1 run as tcgen05 implicit GEMMs on bf16 activations
  bool fuse_resize = true;  // resize + first layer in one kernel (RML_NET_FUSE1=0 disables)
  int R = 0, C = 0, head = 0;
  float alpha = 0.2f;
  std::vector<NetConv> convs;
  int K = 0, act1 = 0, act2 = 0;
  uint16_t* w1t = nullptr;  // bf16 [64][K]
  float *b1 = nullptr, *w2 = nullptr, *b2 = nullptr, *w3 = nullptr, *b3 = nullptr;
  double* kh[3] = {nullptr, nullptr, nullptr};
  double* kv[3] = {nullptr, nullptr, nullptr};
  int2* bh[3] = {nullptr, nullptr, nullptr};
  int2* bv[3] = {nullptr, nullptr, nullptr};
  int ksh[3] = {0, 0, 0}, ksv[3] = {0, 0, 0};
  CUtensorMap map_w1;
  // fused tower kernel (k6_tower.cuh): 0 off, 1 = dnn (resize + layer 1 + layer 2 on chip),
  // 2 = sgan (resize + layer 1 on the tensor cores, NHWC bf16 out)
  int tower_mode = 0;
  uint16_t* t6_w1 = nullptr;   // [3][C1][32] fp16: whi(9), bias hi | whi(9), 0 | wlo(9), bias lo | 0, 0
  float* t6_b1 = nullptr;      // [3][C1]
  int t6_ctas[3] = {0, 0, 0};
  int t6_dbg = 0;              // RML_T6_DBG timing experiments (results are wrong when set)
  bool k4_share = true;        // k4_conv_igemm: kh = 0 / kh = 2 from one activation box (RML_K4_SHARE=0: one box per tap)
};

}  // namespace

struct rml_ctx {
  int device = 0;
  int num_sms = 148;
  int sx = kSX, sy = kSY, sz = kSZ;
  double r_min = 10, r_max = 360, th_min = -42, th_max = 42, ph_min = -30, ph_max = 30;
  float aff_offset = 0.f, aff_scale = 255.f;
  int aff_enabled = 1;
  float* aff_off_dev = nullptr;   // per-feature tables of rml_load_affine (null: scalar pair above)
  float* aff_scl_dev = nullptr;
  int aff_F = 0;
  std::vector<double> aff_shift;  // offset[f] / scale[f]: makes standardised features non-negative (digit path)
  int k1_split = 1, k5_kpg = 0;   // tuning experiments (RML_K1_SPLIT, RML_K5_KPG), read once at create
  int force_f32 = 0;              // rml_set_precision: 1 = float32 features even for an integral model
  int host_narrow = 1;            // rml_predict_host: integral float32 cubes cross the bus as bytes (rml_set_host_narrowing,
                                  // RML_HOST_NARROW=0).  On by default since the results leave through pinned mirrors: the
                                  // conversion of chunk n+1 (65-95 GB/s on the 16-vCPU B200 box) then really overlaps the copy
                                  // and the kernels of chunk n: 113 k -> 140-145 k scans/s end to end
  int host_narrow_threads = 0;    // conversion threads (0 = one per CPU of the process's affinity mask)
  double host_narrow_min_gbs = 57.0;   // break-even against ~53-55 GB/s of PCIe 5 x16 plus the per-chunk launch work
  void* nccl_comm = nullptr;
  int nccl_rank = 0, nccl_world = 1;
  SmallPipe small;
  Model model;
  unsigned int* status = nullptr;  // [0] non-integral values, [1] slice index errors
  int64_t launches = 0;
  std::string err;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  HostPipe pipe;
  Net net;
  // fused K1 || K2 pipeline: projection on most SMs, scorer co-resident on the rest
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_k1a = nullptr, ev_k1b = nullptr, ev_k2b = nullptr;
  unsigned int* tile_done = nullptr;
  int64_t tile_done_cap = 0;
  // uint8 cubes: the projection kernel is issue-bound, not HBM-bound, so it wants every SM (two
  // CTAs each) and the serial K1 -> K2 order wins (profiles/r1c_u8_sweep.txt); k2_sms_u8 > 0
  // (rml_set_fused_u8 / RML_K2_SMS_U8) still selects the co-resident pipeline
  int k2_sms_u8 = 0;
  int k1u8_ctas_per_sm = 2;  // RML_K1U8_CTAS
  int k2_sms = 32;           // SMs reserved for the scorer in fused mode (RML_K2_SMS); set from the SM count in rml_create
  int64_t fused_min_b = 8192;
  int fused_enabled = 1;     // RML_FUSED=0 disables
  int last_fused = 0;
  // scratch of the multi-digit scorer for stand-alone rml_score calls (sized by rml_reserve;
  // rml_predict takes its digit planes from the caller's workspace)
  uint8_t* dg_planes = nullptr;
  long long* dg_norms = nullptr;
  int64_t dg_cap = 0;
  // zoom operators (common.py:143 ndimage.zoom as separable matrices), per projection
  double* zoom_ar[3] = {nullptr, nullptr, nullptr};
  double* zoom_ac[3] = {nullptr, nullptr, nullptr};
  int zoom_ih[3] = {0, 0, 0}, zoom_iw[3] = {0, 0, 0}, zoom_oh[3] = {0, 0, 0}, zoom_ow[3] = {0, 0, 0};
};

namespace {

int fail(rml_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define RML_CUDA(c, expr)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (expr);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail((c), RML_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                 \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

int feature_len(const rml_ctx* c, uint32_t mask) {
  int f = 0;
  if (mask & RML_MASK_XZ) f += c->sx * c->sz;
  if (mask & RML_MASK_YZ) f += c->sy * c->sz;
  if (mask & RML_MASK_XY) f += c->sx * c->sy;
  return f;
}
int feature_stride(const rml_ctx* c, uint32_t mask, int dtype) {
  const int f = feature_len(c, mask);
  return dtype == RML_U8 ? round_up(f, 128) : f;
}

void free_model(Model& m) {
  cudaFree(m.sv_u8);
  cudaFree(m.svnorm);
  cudaFree(m.sv_f64);
  cudaFree(m.coef);
  cudaFree(m.pairw);
  cudaFree(m.sv_digits);
  cudaFree(m.svnorm64);
  cudaFree(m.pairw_dg);
  cudaFree(m.rho);
  cudaFree(m.platt_a);
  cudaFree(m.platt_b);
  cudaFree(m.dg_shift);
  m = Model();
}

template <typename T>
int upload(rml_ctx* c, T** dst, const T* src, size_t n) {
  RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(dst), (n ? n : 1) * sizeof(T)));
  if (n) RML_CUDA(c, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return RML_OK;
}

// u8 row-major [rows][kpad] tensor, box {128 bytes, box_rows}, 128-byte swizzle, OOB -> 0.
int encode_u8_map(rml_ctx* c, CUtensorMap* map, const void* base, int64_t rows, int valid_k,
                  int kpad, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(valid_k), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kpad)};
  cuuint32_t box[2] = {128u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = c->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim,
                         gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, RML_E_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
  return RML_OK;
}

template <int C>
int launch_rbf_i8(rml_ctx* c, const CUtensorMap& map_feats, const K2Params& p, cudaStream_t st,
                  int grid_limit) {
  const int smem = k2_smem_bytes(p.n_tile, p.stages, p.n_pad, C * (C - 1) / 2);
  if (p.stages < 2)   // rml_load_svc_rbf routes such models to the digit / general scorer
    return fail(c, RML_E_UNSUPPORTED, "k2_rbf_i8: %d support vectors x %d classes do not leave room for a 2-stage pipeline", p.n_sv, C);
  const int64_t tiles = (p.B + kK2BlockM - 1) / kK2BlockM;
  const int sms = grid_limit > 0 ? grid_limit : c->num_sms;
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  k2_rbf_i8<C><<<grid, kK2Threads, smem, st>>>(map_feats, c->model.map_sv, p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}
template <int C>
int launch_rbf_digits(rml_ctx* c, const DgMaps& maps, const K2DgParams& p, cudaStream_t st, int grid_limit = 0) {
  const int smem = k2dg_smem_bytes(p.n_pad, C * (C - 1) / 2);
  const int64_t tiles = (p.B + kK2BlockM - 1) / kK2BlockM;
  const int sms = grid_limit > 0 ? grid_limit : c->num_sms;
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  k2_rbf_digits<C><<<grid, kK2Threads, smem, st>>>(maps, p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}
template <int C>
int launch_rbf_general(rml_ctx* c, const K2GenParams& p, cudaStream_t st) {
  const int64_t grid = (p.B + 7) / 8;
  k2_rbf_general<C><<<static_cast<unsigned>(grid), 256, 0, st>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}
template <int C>
int launch_linear(rml_ctx* c, const K2LinParams& p, cudaStream_t st) {
  const int64_t grid = (p.B + 7) / 8;
  k2_linear<C><<<static_cast<unsigned>(grid), 256, 0, st>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

#define DISPATCH_C(Cval, CALL)                                  \
  switch (Cval) {                                               \
    case 2: { constexpr int CC = 2; return CALL; }              \
    case 3: { constexpr int CC = 3; return CALL; }              \
    case 4: { constexpr int CC = 4; return CALL; }              \
    case 5: { constexpr int CC = 5; return CALL; }              \
    case 6: { constexpr int CC = 6; return CALL; }              \
    default: return fail(c, RML_E_UNSUPPORTED, "n_classes=%d not in [2,6]", Cval); \
  }

struct Affine {
  float offset, scale;
  int enabled;
  const float* off_tab = nullptr;   // per-feature tables (device), replace the scalar pair
  const float* scl_tab = nullptr;
};

// every kernel that needs more than 48 KB of dynamic shared memory opts in ONCE, at rml_create
template <typename K>
cudaError_t opt_in_smem(K kernel) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
}
int set_kernel_attributes(rml_ctx* c) {
  RML_CUDA(c, opt_in_smem(k1_project_max<uint8_t>));
  RML_CUDA(c, opt_in_smem(k1_project_max<float>));
  RML_CUDA(c, opt_in_smem(k1_project_max<uint8_t, true>));
  RML_CUDA(c, opt_in_smem(k1_project_max<float, true>));
  RML_CUDA(c, opt_in_smem(k1_project_max_u8in<uint8_t>));
  RML_CUDA(c, opt_in_smem(k1_project_max_u8in<float>));
  RML_CUDA(c, opt_in_smem(k2_rbf_i8<2>)); RML_CUDA(c, opt_in_smem(k2_rbf_i8<3>)); RML_CUDA(c, opt_in_smem(k2_rbf_i8<4>));
  RML_CUDA(c, opt_in_smem(k2_rbf_i8<5>)); RML_CUDA(c, opt_in_smem(k2_rbf_i8<6>));
  RML_CUDA(c, opt_in_smem(k2_rbf_digits<2>)); RML_CUDA(c, opt_in_smem(k2_rbf_digits<3>)); RML_CUDA(c, opt_in_smem(k2_rbf_digits<4>));
  RML_CUDA(c, opt_in_smem(k2_rbf_digits<5>)); RML_CUDA(c, opt_in_smem(k2_rbf_digits<6>));
  RML_CUDA(c, opt_in_smem(k0_derive_targets));
  RML_CUDA(c, opt_in_smem(k0_zoom_concat));
  RML_CUDA(c, opt_in_smem(k3_resize_pil));
  RML_CUDA(c, opt_in_smem(k34_resize_conv1<2>));
  RML_CUDA(c, opt_in_smem(k34_resize_conv1<4>));
  RML_CUDA(c, opt_in_smem(k4_conv3x3s2));
  RML_CUDA(c, opt_in_smem(k4_conv_igemm));
  RML_CUDA(c, opt_in_smem(k5_dense_stack));
  RML_CUDA(c, opt_in_smem(k6_tower<64, true, 1>));
  RML_CUDA(c, opt_in_smem(k6_tower<128, false, 2>));
  return RML_OK;
}

// cubes: float32 voxels (predict.py:91) or, with cube_u8 != 0, the sensor's integers as uint8
// derive != null (fast path only): DerivedTarget's axis sums and top-T indices from the same pass
struct DeriveOut {
  int T;
  int32_t* ijk;
  float* sums;
};
int project_impl(rml_ctx* c, const void* cubes, int64_t B, int mode, const int32_t* ijk,
                 uint32_t mask, int dtype, void* feats, int32_t* norms, cudaStream_t st,
                 int grid_limit = 0, unsigned int* tile_done = nullptr, const Affine* aff_in = nullptr,
                 int cube_u8 = 0, int64_t cube_stride = -1, const DeriveOut* derive = nullptr) {
  const Affine aff = aff_in ? *aff_in : Affine{c->aff_offset, c->aff_scale, c->aff_enabled, c->aff_off_dev, c->aff_scl_dev};
  if (B < 0 || !cubes || !feats) return fail(c, RML_E_INVALID, "rml_project: null buffer or B<0");
  if (aff.off_tab && aff.enabled && dtype == RML_F32 && c->aff_F != feature_len(c, mask))
    return fail(c, RML_E_INVALID, "rml_project: the per-feature affine has F=%d, mask %u gives F=%d",
                c->aff_F, mask, feature_len(c, mask));
  if ((mask & RML_MASK_ALL) == 0 || (mask & ~RML_MASK_ALL))
    return fail(c, RML_E_INVALID, "rml_project: mask %u selects no projection", mask);
  if (mode != RML_MODE_MAX && mode != RML_MODE_SLICE) return fail(c, RML_E_INVALID, "bad mode %d", mode);
  if (mode == RML_MODE_SLICE && !ijk) return fail(c, RML_E_INVALID, "SLICE mode needs ijk");
  if (dtype != RML_F32 && dtype != RML_U8) return fail(c, RML_E_INVALID, "bad dtype %d", dtype);
  const bool fast = mode == RML_MODE_MAX && c->sx == kSX && c->sy == kSY && c->sz == kSZ && cube_stride < 0;
  if (derive && (!fast || cube_u8)) return fail(c, RML_E_UNSUPPORTED, "project_impl: fused derive needs the default arena and float32 cubes");
  // bulk copies / float4 loads need 16-byte granules; uint8 cubes outside the streaming kernel
  // are read as uchar4 at most
  const uintptr_t cube_align = (cube_u8 && !fast) ? 3 : 15;
  if ((reinterpret_cast<uintptr_t>(cubes) & cube_align) || (reinterpret_cast<uintptr_t>(feats) & 15))
    return fail(c, RML_E_INVALID, "cubes/feats must be 16-byte aligned");
  if (B == 0) return RML_OK;
  const int F = feature_len(c, mask);
  const int stride = feature_stride(c, mask, dtype);
  if (fast && cube_u8) {
    K1U8Params p;
    p.cubes = static_cast<const uint8_t*>(cubes); p.feats = feats; p.norms = norms; p.B = B;
    p.stride = stride; p.F = F; p.mask = mask;
    p.offset = aff.offset; p.scale = aff.scale; p.affine = aff.enabled;
    p.aff_off = aff.off_tab; p.aff_scl = aff.scl_tab;
    p.tile_done = dtype == RML_U8 ? tile_done : nullptr;
    const int sms = grid_limit > 0 ? grid_limit : c->num_sms;
    const int64_t ctas = static_cast<int64_t>(sms) *
                         (dtype == RML_U8 && grid_limit <= 0 ? c->k1u8_ctas_per_sm : 1);
    const int grid = static_cast<int>(B < ctas ? B : ctas);
    if (dtype == RML_U8) {
      const int smem = k1u8_smem_bytes<uint8_t>();
      k1_project_max_u8in<uint8_t><<<grid, kK1Threads, smem, st>>>(p);
    } else {
      const int smem = k1u8_smem_bytes<float>();
      k1_project_max_u8in<float><<<grid, kK1Threads, smem, st>>>(p);
    }
  } else if (fast) {
    K1Params p;
    p.cubes = static_cast<const float*>(cubes); p.feats = feats; p.norms = norms; p.status = c->status; p.B = B;
    p.stride = stride; p.F = F; p.mask = mask;
    p.offset = aff.offset; p.scale = aff.scale; p.affine = aff.enabled;
    p.aff_off = aff.off_tab; p.aff_scl = aff.scl_tab;
    p.tile_done = dtype == RML_U8 ? tile_done : nullptr;
    p.split = c->k1_split;
    p.dt_ijk = derive ? derive->ijk : nullptr; p.dt_sums = derive ? derive->sums : nullptr; p.dt_T = derive ? derive->T : 0;
    const int sms = grid_limit > 0 ? grid_limit : c->num_sms;
    const int grid = static_cast<int>(B < sms ? B : sms);
    if (derive) {
      if (dtype == RML_U8) k1_project_max<uint8_t, true><<<grid, kK1Threads, k1_smem_bytes<uint8_t, true>(), st>>>(p);
      else k1_project_max<float, true><<<grid, kK1Threads, k1_smem_bytes<float, true>(), st>>>(p);
    } else if (dtype == RML_U8) {
      const int smem = k1_smem_bytes<uint8_t>();
      k1_project_max<uint8_t><<<grid, kK1Threads, smem, st>>>(p);
    } else {
      const int smem = k1_smem_bytes<float>();
      k1_project_max<float><<<grid, kK1Threads, smem, st>>>(p);
    }
  } else {
    K1GenParams p;
    p.cubes = cubes; p.ijk = ijk; p.feats = feats; p.norms = norms; p.status = c->status; p.B = B;
    p.sx = c->sx; p.sy = c->sy; p.sz = c->sz; p.stride = stride; p.F = F; p.mask = mask;
    p.offset = aff.offset; p.scale = aff.scale; p.affine = aff.enabled; p.mode = mode;
    p.aff_off = aff.off_tab; p.aff_scl = aff.scl_tab;
    p.cube_stride = cube_stride >= 0 ? cube_stride : static_cast<int64_t>(c->sx) * c->sy * c->sz;
    const int64_t want = B < 8ll * c->num_sms ? B : 8ll * c->num_sms;
    const int grid = static_cast<int>(want);
    const bool vec_slice = mode == RML_MODE_SLICE && (c->sz & 3) == 0 && (dtype != RML_U8 || (stride & 3) == 0);
    if (vec_slice) {
      if (cube_u8) {
        if (dtype == RML_U8) k1_project_slice<uint8_t, uint8_t><<<grid, 256, 0, st>>>(p);
        else k1_project_slice<float, uint8_t><<<grid, 256, 0, st>>>(p);
      } else {
        if (dtype == RML_U8) k1_project_slice<uint8_t, float><<<grid, 256, 0, st>>>(p);
        else k1_project_slice<float, float><<<grid, 256, 0, st>>>(p);
      }
    } else {
      if (cube_u8) {
        if (dtype == RML_U8) k1_project_generic<uint8_t, uint8_t><<<grid, 256, 0, st>>>(p);
        else k1_project_generic<float, uint8_t><<<grid, 256, 0, st>>>(p);
      } else {
        if (dtype == RML_U8) k1_project_generic<uint8_t, float><<<grid, 256, 0, st>>>(p);
        else k1_project_generic<float, float><<<grid, 256, 0, st>>>(p);
      }
    }
  }
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

// scratch the multi-digit scorer needs for B float32 feature rows: 3 digit planes + 64-bit norms
size_t digit_scratch_bytes(const rml_ctx* c, int64_t B) {
  const Model& m = c->model;
  if (m.kind != 1 || !m.digits_ok || B <= 0) return 0;
  return align256(static_cast<size_t>(3) * B * m.kpad) + align256(static_cast<size_t>(B) * 8);
}

int score_impl(rml_ctx* c, const void* feats, int dtype, const int32_t* norms, int64_t B,
               double min_proba, float* proba, float* decision, int32_t* label, uint8_t* known,
               cudaStream_t st, int grid_limit = 0, const unsigned int* tile_ready = nullptr,
               void* dg_scratch = nullptr, bool planes_ready = false) {
  Model& m = c->model;
  if (m.kind == 0) return fail(c, RML_E_NOMODEL, "rml_score: no model loaded");
  if (!feats || !proba || !label || B < 0) return fail(c, RML_E_INVALID, "rml_score: null buffer or B<0");
  if (dtype != RML_F32 && dtype != RML_U8 && dtype != RML_F32_EXACT) return fail(c, RML_E_INVALID, "rml_score: bad dtype %d", dtype);
  if (B == 0) return RML_OK;
  if (m.kind == 2) {
    K2LinParams p;
    p.B = B; p.F = m.F; p.dtype = dtype == RML_U8 ? 1 : 0;
    p.stride = dtype == RML_U8 ? round_up(m.F, 128) : m.F;
    p.feats = feats; p.coef = m.coef; p.intercept = m.rho; p.platt_a = m.platt_a; p.platt_b = m.platt_b;
    p.inv_scale = 1.0 / m.feature_scale; p.feature_scale = m.feature_scale; p.min_proba = min_proba;
    p.proba = proba; p.decision = decision; p.label = label; p.known = known;
    DISPATCH_C(m.C, launch_linear<CC>(c, p, st));
  }
  if (dtype == RML_U8) {
    if (!m.integral)
      return fail(c, RML_E_UNSUPPORTED,
                  "rml_score: u8 features need an integral SVC model (support vectors are not "
                  "integers/%g); project to RML_F32 instead", m.feature_scale);
    if (!norms) return fail(c, RML_E_INVALID, "rml_score: u8 features need norms_dev");
    if (reinterpret_cast<uintptr_t>(feats) & 15) return fail(c, RML_E_INVALID, "feats must be 16-byte aligned");
    CUtensorMap map_feats;
    int rc = encode_u8_map(c, &map_feats, feats, B, m.F, m.kpad, kK2BlockM);
    if (rc) return rc;
    K2Params p;
    p.B = B; p.n_sv = m.n_sv; p.n_tile = m.n_tile; p.n_chunks = m.n_chunks; p.k_blocks = m.kpad / 128;
    p.n_pad = m.n_tile * m.n_chunks;
    p.stages = m.i8_stages;
    p.unorm = norms; p.svnorm = m.svnorm; p.pairw = m.pairw; p.rho = m.rho;
    p.platt_a = m.platt_a; p.platt_b = m.platt_b;
    p.neg_gamma_s2 = -m.gamma / (m.feature_scale * m.feature_scale);
    p.min_proba = min_proba; p.proba = proba; p.decision = decision; p.label = label; p.known = known;
    p.tile_ready = tile_ready;
    DISPATCH_C(m.C, launch_rbf_i8<CC>(c, map_feats, p, st, grid_limit));
  }
  if (dtype == RML_F32 && m.digits_ok) {
    // tensor-core exact path: float32 features -> 24-bit fixed point digit planes -> 9 u8 GEMMs.
    // The planes live in caller-provided scratch (rml_predict's workspace) or in the context
    // scratch sized by rml_reserve; nothing is allocated here.
    if (reinterpret_cast<uintptr_t>(feats) & 3) return fail(c, RML_E_INVALID, "feats must be 4-byte aligned");
    uint8_t* planes;
    long long* dnorms;
    if (dg_scratch) {
      planes = static_cast<uint8_t*>(dg_scratch);
      dnorms = reinterpret_cast<long long*>(planes + align256(static_cast<size_t>(3) * B * m.kpad));
    } else {
      if (c->dg_cap < B)
        return fail(c, RML_E_INVALID, "rml_score: float32 features of this model need %lld rows of digit scratch, "
                    "%lld reserved; call rml_reserve(ctx, max_batch, RML_RESERVE_SCORE) first",
                    static_cast<long long>(B), static_cast<long long>(c->dg_cap));
      planes = c->dg_planes;
      dnorms = c->dg_norms;
    }
    if (!planes_ready) {
      DgQuantParams qp;
      qp.feats = static_cast<const float*>(feats); qp.planes = planes; qp.norms = dnorms;
      qp.status = c->status; qp.B = B; qp.F = m.F; qp.stride = m.kpad; qp.scale = m.dg_scale;
      qp.shift = m.dg_shift;
      k1_quantize_digits<<<static_cast<unsigned>((B + 7) / 8), 256, 0, st>>>(qp);
      RML_CUDA(c, cudaGetLastError());
      ++c->launches;
    }
    DgMaps maps;
    const size_t plane = static_cast<size_t>(B) * m.kpad;
    for (int d = 0; d < 3; ++d) {
      int rc = encode_u8_map(c, &maps.a[d], planes + d * plane, B, m.F, m.kpad, kK2BlockM);
      if (rc) return rc;
      maps.b[d] = m.map_sv_dg[d];
    }
    K2DgParams dp;
    dp.B = B; dp.n_sv = m.n_sv; dp.n_chunks = m.dg_chunks; dp.k_blocks = m.kpad / 128; dp.n_pad = m.dg_pad;
    dp.unorm = dnorms; dp.svnorm = m.svnorm64; dp.pairw = m.pairw_dg; dp.rho = m.rho;
    dp.platt_a = m.platt_a; dp.platt_b = m.platt_b;
    dp.tile_ready = tile_ready;
    dp.neg_gamma_fixed = -m.gamma / (m.dg_scale * m.dg_scale * 4294967296.0);
    dp.min_proba = min_proba; dp.proba = proba; dp.decision = decision; dp.label = label; dp.known = known;
    DISPATCH_C(m.C, launch_rbf_digits<CC>(c, maps, dp, st, grid_limit));
  }
  K2GenParams p;
  p.B = B; p.F = m.F; p.n_sv = m.n_sv; p.feats = static_cast<const float*>(feats); p.sv = m.sv_f64;
  p.coef = m.coef; p.rho = m.rho; p.platt_a = m.platt_a; p.platt_b = m.platt_b;
  p.neg_gamma = -m.gamma; p.min_proba = min_proba;
  p.proba = proba; p.decision = decision; p.label = label; p.known = known;
  for (int i = 0; i < kMaxClasses; ++i) p.class_end[i] = m.class_end[i];
  DISPATCH_C(m.C, launch_rbf_general<CC>(c, p, st));
}

// u8 operand rows + the integer tensor-core scorer: an integral SVC (or a linear model, whose
// scorer reads either row type) AND no per-feature affine AND not overridden by rml_set_precision
bool use_u8_path(const rml_ctx* c) {
  if (c->force_f32 || c->aff_off_dev) return false;
  return c->model.kind == 2 || (c->model.kind == 1 && c->model.integral);
}

// Workspace plan of rml_predict for B scans (everything the call needs; nothing is allocated):
//   [feature rows][norms int32][digit planes + 64-bit norms (float32 path of an SVC)][tile counters]
struct PredictPlan {
  int dtype;
  size_t feat_bytes, norm_off, dg_off, tile_off, total;
};
PredictPlan predict_plan(const rml_ctx* c, int64_t B, uint32_t mask) {
  PredictPlan pl;
  pl.dtype = use_u8_path(c) ? RML_U8 : RML_F32;
  const size_t stride = feature_stride(c, mask, pl.dtype);
  pl.feat_bytes = align256(static_cast<size_t>(B) * stride * (pl.dtype == RML_U8 ? 1 : 4));
  pl.norm_off = pl.feat_bytes;
  pl.dg_off = pl.norm_off + align256(static_cast<size_t>(B) * 4);
  pl.tile_off = pl.dg_off + (pl.dtype == RML_F32 ? digit_scratch_bytes(c, B) : 0);
  pl.total = pl.tile_off + align256(static_cast<size_t>((B + kK2BlockM - 1) / kK2BlockM) * 4) + 256;
  return pl;
}

int predict_impl(rml_ctx* c, const void* cubes, int64_t B, int mode, const int32_t* ijk,
                 uint32_t mask, double min_proba, void* work, float* proba, int32_t* label,
                 uint8_t* known, cudaStream_t st, int cube_u8 = 0, int64_t cube_stride = -1) {
  if (c->model.kind == 0) return fail(c, RML_E_NOMODEL, "rml_predict: no model loaded");
  if (feature_len(c, mask) != c->model.F)
    return fail(c, RML_E_INVALID, "rml_predict: mask gives F=%d but the model has F=%d",
                feature_len(c, mask), c->model.F);
  if (!work) return fail(c, RML_E_INVALID, "rml_predict: workspace is null");
  if (reinterpret_cast<uintptr_t>(work) & 255) return fail(c, RML_E_INVALID, "rml_predict: workspace must be 256-byte aligned");
  const PredictPlan pl = predict_plan(c, B, mask);
  const int dtype = pl.dtype;
  char* ws = static_cast<char*>(work);
  int32_t* norms = reinterpret_cast<int32_t*>(ws + pl.norm_off);
  void* dg_scratch = (dtype == RML_F32 && c->model.kind == 1 && c->model.digits_ok) ? ws + pl.dg_off : nullptr;
  unsigned int* tile_done = reinterpret_cast<unsigned int*>(ws + pl.tile_off);
  // the scorer expects features scaled like common.process_samples(scale=True) — or, with a
  // per-feature affine loaded, standardised like the scaler the model was trained behind
  const Affine aff = c->aff_off_dev ? Affine{0.f, 1.f, 1, c->aff_off_dev, c->aff_scl_dev}
                                    : Affine{0.f, static_cast<float>(c->model.feature_scale), 1};
  // uint8 cubes stream 4x faster, so the scorer gets a larger share of the SMs
  const int k2_sms = cube_u8 ? c->k2_sms_u8 : c->k2_sms;
  const bool fused = c->fused_enabled && c->model.kind == 1 && dtype == RML_U8 && mode == RML_MODE_MAX &&
                     c->sx == kSX && c->sy == kSY && c->sz == kSZ && B >= c->fused_min_b &&
                     k2_sms > 0 && k2_sms < c->num_sms && cube_stride < 0;
  c->last_fused = fused ? 1 : 0;
  if (fused) {
    // One pipeline, two co-resident kernels: K1 streams cubes on (num_sms - k2_sms) SMs and
    // bumps a per-tile counter for every finished scan; K2 runs on the remaining SMs and
    // starts on a 128-scan tile as soon as its counter is full (device-side flags, no host sync).
    const int64_t tiles = (B + kK2BlockM - 1) / kK2BlockM;
    // a previous fused call (possibly on another stream) may still be reading its counters
    RML_CUDA(c, cudaStreamWaitEvent(st, c->ev_join, 0));
    RML_CUDA(c, cudaMemsetAsync(tile_done, 0, tiles * sizeof(unsigned int), st));
    RML_CUDA(c, cudaEventRecord(c->ev_fork, st));
    RML_CUDA(c, cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
    if (c->ev_k1a) RML_CUDA(c, cudaEventRecord(c->ev_k1a, st));
    int rc = project_impl(c, cubes, B, mode, ijk, mask, dtype, work, norms, st,
                          c->num_sms - k2_sms, tile_done, &aff, cube_u8);
    if (c->ev_k1b) RML_CUDA(c, cudaEventRecord(c->ev_k1b, st));
    if (rc) return rc;
    rc = score_impl(c, work, dtype, norms, B, min_proba, proba, nullptr, label, known, c->aux_stream,
                    k2_sms, tile_done);
    if (rc) return rc;
    RML_CUDA(c, cudaEventRecord(c->ev_join, c->aux_stream));
    RML_CUDA(c, cudaStreamWaitEvent(st, c->ev_join, 0));
    if (c->ev_k2b) RML_CUDA(c, cudaEventRecord(c->ev_k2b, st));
    return RML_OK;
  }
  if (c->ev_k1a) RML_CUDA(c, cudaEventRecord(c->ev_k1a, st));
  int rc = project_impl(c, cubes, B, mode, ijk, mask, dtype, work, norms, st, 0, nullptr, &aff, cube_u8,
                        cube_stride);
  if (c->ev_k1b) RML_CUDA(c, cudaEventRecord(c->ev_k1b, st));
  if (rc) return rc;
  rc = score_impl(c, work, dtype, norms, B, min_proba, proba, nullptr, label, known, st, 0, nullptr, dg_scratch);
  if (c->ev_k2b && rc == RML_OK) RML_CUDA(c, cudaEventRecord(c->ev_k2b, st));
  return rc;
}

void free_net(Net& n) {
  for (auto& cv : n.convs) {
    for (int b = 0; b < 3; ++b) { cudaFree(cv.w[b]); cudaFree(cv.bias[b]); }
    cudaFree(cv.wt_bf16); cudaFree(cv.bias3);
  }
  cudaFree(n.w1t); cudaFree(n.b1); cudaFree(n.w2); cudaFree(n.b2); cudaFree(n.w3); cudaFree(n.b3);
  cudaFree(n.t6_w1); cudaFree(n.t6_b1);
  for (int b = 0; b < 3; ++b) { cudaFree(n.kh[b]); cudaFree(n.kv[b]); cudaFree(n.bh[b]); cudaFree(n.bv[b]); }
  n = Net();
}

void free_small(SmallPipe& sp) {
  cudaFree(sp.cube); cudaFree(sp.ijk); cudaFree(sp.feats); cudaFree(sp.work);
  cudaFree(sp.proba); cudaFree(sp.label); cudaFree(sp.known);
  if (sp.pinned) cudaFreeHost(sp.pinned);
  sp = SmallPipe();
}

void free_pipe(HostPipe& hp) {
  for (int i = 0; i < kHostBufs; ++i) {
    cudaFree(hp.cubes[i]); cudaFree(hp.ijk[i]); cudaFree(hp.work[i]);
    cudaFree(hp.proba[i]); cudaFree(hp.label[i]); cudaFree(hp.known[i]);
    if (hp.stream[i]) cudaStreamDestroy(hp.stream[i]);
    if (hp.res_pin[i]) cudaFreeHost(hp.res_pin[i]);
    if (hp.res_ev[i]) cudaEventDestroy(hp.res_ev[i]);
    if (hp.narrow_pin[i]) cudaFreeHost(hp.narrow_pin[i]);
    if (hp.narrow_ev[i]) cudaEventDestroy(hp.narrow_ev[i]);
  }
  rml_host::narrow_pool_destroy(hp.pool);
  hp = HostPipe();
}

}  // namespace

// ============================================================================== C ABI
extern "C" {

int rml_version(void) { return 100; }

const char* rml_last_error(const rml_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int rml_create(int device, rml_ctx** out) {
  if (!out) return fail(nullptr, RML_E_INVALID, "rml_create: out is null");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, RML_E_CUDA, "rml_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, RML_E_INVALID, "rml_create: device %d of %d", device, n);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, RML_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, RML_E_CUDA, "rml_create: device %d is sm_%d%d; this build is sm_100a only",
                device, prop.major, prop.minor);
  rml_ctx* c = new rml_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  DeviceGuard g(device);
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete c;
    return fail(nullptr, RML_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  }
  c->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  if (cudaMalloc(reinterpret_cast<void**>(&c->status), 16) != cudaSuccess ||
      cudaMemset(c->status, 0, 16) != cudaSuccess) {
    delete c;
    return fail(nullptr, RML_E_CUDA, "status allocation failed");
  }
  // everything a hot entry point would otherwise create lazily: kernel attributes, the second
  // stream of the co-resident pipeline and its events
  if (set_kernel_attributes(c) != RML_OK ||
      cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    g_create_error = "rml_create: kernel attribute / stream setup failed: " + c->err;
    cudaFree(c->status);
    delete c;
    return RML_E_CUDA;
  }
  if (const char* e6 = getenv("RML_K1_SPLIT")) { const int v = atoi(e6); if (v == 2 || v == 4) c->k1_split = v; }
  if (const char* e7 = getenv("RML_K5_KPG")) { const int v = atoi(e7); if (v == 1 || v == 2 || v == 4) c->k5_kpg = v; }
  // measured optimum on B200 (148 SMs): 32 scorer SMs, 116 projection SMs (profiles/r1_fused_sweep.txt)
  c->k2_sms = (c->num_sms * 32 + 74) / 148;
  if (const char* e1 = getenv("RML_K2_SMS")) c->k2_sms = atoi(e1);
  if (const char* e4 = getenv("RML_K2_SMS_U8")) c->k2_sms_u8 = atoi(e4);
  if (const char* e5 = getenv("RML_K1U8_CTAS")) { const int v = atoi(e5); if (v == 1 || v == 2) c->k1u8_ctas_per_sm = v; }
  if (const char* e2 = getenv("RML_FUSED")) c->fused_enabled = atoi(e2);
  if (const char* e3 = getenv("RML_FUSED_MIN_B")) c->fused_min_b = atoll(e3);
  if (const char* e6 = getenv("RML_HOST_NARROW")) c->host_narrow = atoi(e6) != 0;
  if (const char* e7 = getenv("RML_HOST_NARROW_THREADS")) c->host_narrow_threads = atoi(e7);
  *out = c;
  return RML_OK;
}

int rml_destroy(rml_ctx* c) {
  if (!c) return RML_OK;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  free_model(c->model);
  free_pipe(c->pipe);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  for (cudaEvent_t e : {c->ev_fork, c->ev_join, c->ev_k1a, c->ev_k1b, c->ev_k2b})
    if (e) cudaEventDestroy(e);
  cudaFree(c->tile_done);
  cudaFree(c->dg_planes);
  cudaFree(c->dg_norms);
  cudaFree(c->aff_off_dev);
  cudaFree(c->aff_scl_dev);
  free_small(c->small);
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
  free_net(c->net);
  for (int q = 0; q < 3; ++q) { cudaFree(c->zoom_ar[q]); cudaFree(c->zoom_ac[q]); }
  cudaFree(c->status);
  delete c;
  return RML_OK;
}

int rml_set_arena(rml_ctx* c, int sx, int sy, int sz) {
  if (!c) return RML_E_INVALID;
  if (sx <= 0 || sy <= 0 || sz <= 0) return fail(c, RML_E_INVALID, "arena dims must be positive");
  c->sx = sx; c->sy = sy; c->sz = sz;
  return RML_OK;
}
int rml_set_arena_bounds(rml_ctx* c, double r_min, double r_max, double th_min, double th_max,
                         double ph_min, double ph_max) {
  if (!c) return RML_E_INVALID;
  c->r_min = r_min; c->r_max = r_max; c->th_min = th_min; c->th_max = th_max;
  c->ph_min = ph_min; c->ph_max = ph_max;
  return RML_OK;
}
int rml_feature_len(const rml_ctx* c, uint32_t mask) { return c ? feature_len(c, mask) : RML_E_INVALID; }
int rml_feature_stride(const rml_ctx* c, uint32_t mask, int dtype) {
  return c ? feature_stride(c, mask, dtype) : RML_E_INVALID;
}
int rml_set_affine(rml_ctx* c, float offset, float scale, int enabled) {
  if (!c) return RML_E_INVALID;
  if (enabled && scale == 0.f) return fail(c, RML_E_INVALID, "affine scale must be non-zero");
  if (c->aff_off_dev)
    return fail(c, RML_E_INVALID, "rml_set_affine: a per-feature affine is loaded; clear it with rml_load_affine(ctx, NULL, NULL, 0) first");
  c->aff_offset = offset; c->aff_scale = scale; c->aff_enabled = enabled;
  return RML_OK;
}

int rml_load_affine(rml_ctx* c, const float* offset, const float* scale, int F) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  cudaFree(c->aff_off_dev); cudaFree(c->aff_scl_dev);
  c->aff_off_dev = c->aff_scl_dev = nullptr;
  c->aff_F = 0;
  c->aff_shift.clear();
  // a model digitised under another affine no longer matches its features: it keeps working
  // through the float64 scorer until it is loaded again
  if (c->model.kind == 1) c->model.digits_ok = false;
  if (!offset && !scale) {           // back to the reference default, common.py:148
    c->aff_offset = 0.f; c->aff_scale = 255.f; c->aff_enabled = 1;
    return RML_OK;
  }
  if (!offset || !scale || F <= 0) return fail(c, RML_E_INVALID, "rml_load_affine: offset and scale must both be given with F > 0");
  for (int f = 0; f < F; ++f)
    if (!(scale[f] != 0.f) || !std::isfinite(scale[f]) || !std::isfinite(offset[f]))
      return fail(c, RML_E_INVALID, "rml_load_affine: scale[%d] = %g, offset[%d] = %g", f, scale[f], f, offset[f]);
  int rc;
  if ((rc = upload(c, &c->aff_off_dev, offset, F))) return rc;
  if ((rc = upload(c, &c->aff_scl_dev, scale, F))) return rc;
  c->aff_F = F;
  c->aff_enabled = 1;
  c->aff_shift.resize(F);
  for (int f = 0; f < F; ++f) c->aff_shift[f] = static_cast<double>(offset[f]) / static_cast<double>(scale[f]);
  return RML_OK;
}

int rml_set_precision(rml_ctx* c, int force_f32) {
  if (!c) return RML_E_INVALID;
  c->force_f32 = force_f32 ? 1 : 0;
  return RML_OK;
}

int rml_load_svc_rbf(rml_ctx* c, int C, int F, int n_sv, const int32_t* n_support,
                     const double* sv, const double* dual_coef, const double* rho, double gamma,
                     const double* platt_a, const double* platt_b, double feature_scale) {
  if (!c) return RML_E_INVALID;
  if (C < 2 || C > kMaxClasses || F <= 0 || n_sv <= 0 || !n_support || !sv || !dual_coef || !rho ||
      !platt_a || !platt_b || !(feature_scale > 0))
    return fail(c, RML_E_INVALID, "rml_load_svc_rbf: bad arguments");
  int tot = 0;
  for (int i = 0; i < C; ++i) tot += n_support[i];
  if (tot != n_sv) return fail(c, RML_E_INVALID, "sum(n_support)=%d != n_sv=%d", tot, n_sv);
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  free_model(c->model);
  Model& m = c->model;
  m.C = C; m.F = F; m.n_sv = n_sv; m.gamma = gamma; m.feature_scale = feature_scale;
  int acc = 0;
  for (int i = 0; i < kMaxClasses; ++i) {
    if (i < C) acc += n_support[i];
    m.class_end[i] = acc;
  }
  // integrality: every support-vector component is (integer in [0,255]) / feature_scale
  m.kpad = round_up(F, 128);
  m.n_chunks = (n_sv + kK2MaxTileN - 1) / kK2MaxTileN;
  m.n_tile = round_up((n_sv + m.n_chunks - 1) / m.n_chunks, 16);
  const int n_pad = m.n_tile * m.n_chunks;
  std::vector<uint8_t> u8(static_cast<size_t>(n_pad) * m.kpad, 0);
  std::vector<int32_t> norm(n_pad, 0);
  // a model behind a per-feature affine lives in standardised space: no integer form
  const bool per_feature = c->aff_off_dev != nullptr;
  if (per_feature && c->aff_F != F)
    return fail(c, RML_E_INVALID, "rml_load_svc_rbf: the loaded per-feature affine has F=%d, the model F=%d", c->aff_F, F);
  bool integral = !per_feature;
  for (int n = 0; n < n_sv && integral; ++n) {
    int64_t s2 = 0;
    for (int f = 0; f < F; ++f) {
      const double v = sv[static_cast<size_t>(n) * F + f] * feature_scale;
      const double r = std::nearbyint(v);
      // float32(u / scale) * scale is within 2e-5 of u for every u in [0,255]; anything further
      // from an integer is a genuinely non-integral support vector (augmented data)
      if (std::fabs(v - r) > 2e-5 || r < 0 || r > 255) { integral = false; break; }
      u8[static_cast<size_t>(n) * m.kpad + f] = static_cast<uint8_t>(r);
      s2 += static_cast<int64_t>(r) * static_cast<int64_t>(r);
    }
    norm[n] = static_cast<int32_t>(s2);
  }
  // 2*max(u.s) and the norm sum must stay inside s32: F * 255^2 * 2 < 2^31  <=>  F <= 16512
  if (static_cast<double>(F) * 65025.0 * 2.0 >= 2147483647.0) integral = false;
  // many support vectors x many classes can leave no room for a 2-stage operand ring next to the
  // epilogue tables: such a model takes the digit / general scorer instead of failing at predict
  m.i8_stages = k2_pick_stages(m.n_tile, n_pad, C * (C - 1) / 2);
  if (m.i8_stages < 2) integral = false;
  m.integral = integral;
  int rc;
  if (integral) {
    // weight of SV n (class cls) in OvO pair (i<j): dual_coef[j-1][n] if cls == i, dual_coef[i][n]
    // if cls == j (svm.cpp:2864-2884), 0 otherwise; padded columns stay 0
    const int NP = C * (C - 1) / 2;
    std::vector<double> pw(static_cast<size_t>(NP) * n_pad, 0.0);
    for (int n = 0, cls = 0; n < n_sv; ++n) {
      while (n >= m.class_end[cls]) ++cls;
      int pidx = 0;
      for (int i = 0; i < C; ++i)
        for (int j = i + 1; j < C; ++j, ++pidx) {
          if (cls == i) pw[static_cast<size_t>(pidx) * n_pad + n] = dual_coef[static_cast<size_t>(j - 1) * n_sv + n];
          else if (cls == j) pw[static_cast<size_t>(pidx) * n_pad + n] = dual_coef[static_cast<size_t>(i) * n_sv + n];
        }
    }
    if ((rc = upload(c, &m.pairw, pw.data(), pw.size()))) return rc;
    if ((rc = upload(c, &m.sv_u8, u8.data(), u8.size()))) return rc;
    if ((rc = upload(c, &m.svnorm, norm.data(), norm.size()))) return rc;
    if ((rc = encode_u8_map(c, &m.map_sv, m.sv_u8, n_sv, F, m.kpad, m.n_tile))) return rc;
  }
  // digit planes: every (shifted) component in [0, 256/dg_scale) and room for 3 products per s32
  // accumulator.  Without a per-feature affine dg_scale = feature_scale and the shift is 0; with
  // one (features standardised as (u - offset_f)/scale_f) the shift offset_f/scale_f makes every
  // component non-negative — distances are translation invariant — and dg_scale is chosen so that
  // 255/min|scale_f| still fits 24 bits.
  {
    bool ok = static_cast<double>(F) * 65025.0 * 3.0 < 2147483647.0;
    m.dg_chunks = (n_sv + kDgTileN - 1) / kDgTileN;
    m.dg_pad = m.dg_chunks * kDgTileN;
    const int NP = C * (C - 1) / 2;
    if (k2dg_smem_bytes(m.dg_pad, NP) > 232448) ok = false;
    m.dg_scale = feature_scale;
    if (per_feature) {
      double top = 0.0;
      // largest standardised-and-shifted value a sensor byte can take: 255 / scale_f
      std::vector<float> scl(F);
      cudaMemcpy(scl.data(), c->aff_scl_dev, static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost);
      for (int f = 0; f < F; ++f) {
        if (scl[f] < 0.f) { ok = false; break; }      // a negative scale flips the axis; not a fitted scaler
        const double v = 255.0 / static_cast<double>(scl[f]);
        if (v > top) top = v;
      }
      for (int n = 0; n < n_sv && ok; ++n)
        for (int f = 0; f < F; ++f) {
          const double v = sv[static_cast<size_t>(n) * F + f] + c->aff_shift[f];
          if (v > top) top = v;
        }
      m.dg_scale = top > 0.0 ? 255.5 / top : 1.0;
    }
    std::vector<uint8_t> dg;
    std::vector<long long> n64(m.dg_pad, 0);
    if (ok) dg.assign(static_cast<size_t>(3) * m.dg_pad * m.kpad, 0);
    const size_t plane = static_cast<size_t>(m.dg_pad) * m.kpad;
    for (int n = 0; n < n_sv && ok; ++n) {
      unsigned long long s2 = 0;
      for (int f = 0; f < F; ++f) {
        const double x = sv[static_cast<size_t>(n) * F + f] + (per_feature ? c->aff_shift[f] : 0.0);
        const double v = std::nearbyint(x * m.dg_scale * 65536.0);
        if (!(v >= 0.0 && v < 16777216.0)) { ok = false; break; }
        const uint32_t X = static_cast<uint32_t>(v);
        const size_t o = static_cast<size_t>(n) * m.kpad + f;
        dg[o] = static_cast<uint8_t>(X & 255u);
        dg[plane + o] = static_cast<uint8_t>((X >> 8) & 255u);
        dg[2 * plane + o] = static_cast<uint8_t>(X >> 16);
        s2 += static_cast<unsigned long long>(X) * X;
      }
      n64[n] = static_cast<long long>(s2);
    }
    m.digits_ok = ok;
    if (ok) {
      std::vector<double> pw(static_cast<size_t>(NP) * m.dg_pad, 0.0);
      for (int n = 0, cls = 0; n < n_sv; ++n) {
        while (n >= m.class_end[cls]) ++cls;
        int pidx = 0;
        for (int i = 0; i < C; ++i)
          for (int j = i + 1; j < C; ++j, ++pidx) {
            if (cls == i) pw[static_cast<size_t>(pidx) * m.dg_pad + n] = dual_coef[static_cast<size_t>(j - 1) * n_sv + n];
            else if (cls == j) pw[static_cast<size_t>(pidx) * m.dg_pad + n] = dual_coef[static_cast<size_t>(i) * n_sv + n];
          }
      }
      if ((rc = upload(c, &m.sv_digits, dg.data(), dg.size()))) return rc;
      if ((rc = upload(c, &m.svnorm64, n64.data(), n64.size()))) return rc;
      if ((rc = upload(c, &m.pairw_dg, pw.data(), pw.size()))) return rc;
      if (per_feature && (rc = upload(c, &m.dg_shift, c->aff_shift.data(), static_cast<size_t>(F)))) return rc;
      for (int d = 0; d < 3; ++d)
        if ((rc = encode_u8_map(c, &m.map_sv_dg[d], m.sv_digits + d * plane, n_sv, F, m.kpad, kDgTileN))) return rc;
    }
  }
  if ((rc = upload(c, &m.sv_f64, sv, static_cast<size_t>(n_sv) * F))) return rc;
  if ((rc = upload(c, &m.coef, dual_coef, static_cast<size_t>(C - 1) * n_sv))) return rc;
  if ((rc = upload(c, &m.rho, rho, static_cast<size_t>(C) * (C - 1) / 2))) return rc;
  const int ncal = C == 2 ? 1 : C;
  if ((rc = upload(c, &m.platt_a, platt_a, ncal))) return rc;
  if ((rc = upload(c, &m.platt_b, platt_b, ncal))) return rc;
  m.kind = 1;
  return RML_OK;
}

int rml_load_linear(rml_ctx* c, int C, int F, const double* coef, const double* intercept,
                    const double* platt_a, const double* platt_b, double feature_scale) {
  if (!c) return RML_E_INVALID;
  if (C < 2 || C > kMaxClasses || F <= 0 || !coef || !intercept || !platt_a || !platt_b ||
      !(feature_scale > 0))
    return fail(c, RML_E_INVALID, "rml_load_linear: bad arguments");
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  free_model(c->model);
  Model& m = c->model;
  m.C = C; m.F = F; m.feature_scale = feature_scale; m.integral = true;
  const int R = C == 2 ? 1 : C;
  int rc;
  if ((rc = upload(c, &m.coef, coef, static_cast<size_t>(R) * F))) return rc;
  if ((rc = upload(c, &m.rho, intercept, R))) return rc;
  if ((rc = upload(c, &m.platt_a, platt_a, R))) return rc;
  if ((rc = upload(c, &m.platt_b, platt_b, R))) return rc;
  m.kind = 2;
  return RML_OK;
}

int rml_model_is_integral(const rml_ctx* c) { return c && c->model.kind && c->model.integral; }

int rml_project(rml_ctx* c, const float* cubes, int64_t B, int mode, const int32_t* ijk,
                uint32_t mask, int dtype, void* feats, int32_t* norms, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  return project_impl(c, cubes, B, mode, ijk, mask, dtype, feats, norms, static_cast<cudaStream_t>(stream));
}

int rml_project_u8(rml_ctx* c, const uint8_t* cubes, int64_t B, int mode, const int32_t* ijk,
                   uint32_t mask, int dtype, void* feats, int32_t* norms, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  return project_impl(c, cubes, B, mode, ijk, mask, dtype, feats, norms, static_cast<cudaStream_t>(stream),
                      0, nullptr, nullptr, 1);
}

int rml_process_samples(rml_ctx* c, const float* xz, const float* yz, const float* xy, int64_t B,
                        uint32_t mask, int scale, float* feats, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!feats || B < 0) return fail(c, RML_E_INVALID, "rml_process_samples: null output or B<0");
  if ((mask & RML_MASK_ALL) == 0) return fail(c, RML_E_INVALID, "mask selects no projection");
  if (((mask & 1) && !xz) || ((mask & 2) && !yz) || ((mask & 4) && !xy))
    return fail(c, RML_E_INVALID, "rml_process_samples: a selected projection is null");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  PsParams p;
  p.proj[0] = (mask & 1) ? xz : nullptr; p.proj[1] = (mask & 2) ? yz : nullptr; p.proj[2] = (mask & 4) ? xy : nullptr;
  p.len[0] = c->sx * c->sz; p.len[1] = c->sy * c->sz; p.len[2] = c->sx * c->sy;
  p.off[0] = 0;
  p.off[1] = (mask & 1) ? p.len[0] : 0;
  p.off[2] = p.off[1] + ((mask & 2) ? p.len[1] : 0);
  p.feats = feats; p.B = B; p.F = feature_len(c, mask); p.scale = scale;
  // scale != 0 applies the context affine: default (0, 255) = common.py:148; (127.5, 127.5) = dnn.py:203
  p.offset = c->aff_offset; p.scale_value = c->aff_scale;
  p.aff_off = c->aff_off_dev; p.aff_scl = c->aff_scl_dev;
  if (scale && c->aff_off_dev && c->aff_F != p.F)
    return fail(c, RML_E_INVALID, "rml_process_samples: the per-feature affine has F=%d, mask %u gives F=%d", c->aff_F, mask, p.F);
  const int64_t total = B * p.F;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 16ll * c->num_sms) blocks = 16ll * c->num_sms;
  k1_process_samples<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

int rml_matrix_indices(rml_ctx* c, const double* xyz, int64_t B, int32_t* ijk, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!xyz || !ijk || B < 0) return fail(c, RML_E_INVALID, "rml_matrix_indices: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  IdxParams p;
  p.xyz = xyz; p.ijk = ijk; p.B = B; p.sx = c->sx; p.sy = c->sy; p.sz = c->sz;
  p.r_min = c->r_min; p.r_max = c->r_max; p.th_min = c->th_min; p.th_max = c->th_max;
  p.ph_min = c->ph_min; p.ph_max = c->ph_max;
  k1_matrix_indices<<<static_cast<unsigned>((B + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

int rml_quantize_features(rml_ctx* c, const float* feats, int64_t B, int F, uint8_t* out,
                          int32_t* norms, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!feats || !out || !norms || B < 0 || F <= 0)
    return fail(c, RML_E_INVALID, "rml_quantize_features: null buffer or bad shape");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  QuantParams p;
  p.feats = feats; p.out = out; p.norms = norms; p.status = c->status; p.B = B; p.F = F;
  p.stride = round_up(F, 128);
  p.scale = static_cast<float>(c->model.kind ? c->model.feature_scale : 255.0);
  k1_quantize<<<static_cast<unsigned>((B + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

int rml_score(rml_ctx* c, const void* feats, int dtype, const int32_t* norms, int64_t B,
              double min_proba, float* proba, float* decision, int32_t* label, uint8_t* known,
              rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  return score_impl(c, feats, dtype, norms, B, min_proba, proba, decision, label, known,
                    static_cast<cudaStream_t>(stream));
}

size_t rml_predict_workspace_bytes(const rml_ctx* c, int64_t B) {
  if (!c || B <= 0) return 256;
  return predict_plan(c, B, RML_MASK_ALL).total;
}

int rml_predict(rml_ctx* c, const float* cubes, int64_t B, int mode, const int32_t* ijk,
                uint32_t mask, double min_proba, void* work, float* proba, int32_t* label,
                uint8_t* known, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  return predict_impl(c, cubes, B, mode, ijk, mask, min_proba, work, proba, label, known,
                      static_cast<cudaStream_t>(stream));
}

int rml_predict_u8(rml_ctx* c, const uint8_t* cubes, int64_t B, int mode, const int32_t* ijk,
                   uint32_t mask, double min_proba, void* work, float* proba, int32_t* label,
                   uint8_t* known, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  return predict_impl(c, cubes, B, mode, ijk, mask, min_proba, work, proba, label, known,
                      static_cast<cudaStream_t>(stream), 1);
}

}  // extern "C"

namespace {
int reserve_host_pipe(rml_ctx* c, int cube_u8) {
  HostPipe& hp = c->pipe;
  // 512 float32 cubes = 246 MB per H2D transfer; uint8 cubes are a quarter of that, so twice the
  // scans per chunk keeps the transfers long and the launches few
  const int64_t chunk = cube_u8 ? 1024 : 512;
  const size_t cube_bytes = static_cast<size_t>(c->sx) * c->sy * c->sz * (cube_u8 ? 1 : 4);
  // the feature staging depends on the loaded model (u8 rows vs float32 rows + digit planes)
  const size_t work_need = rml_predict_workspace_bytes(c, chunk);
  if (hp.chunk == chunk && hp.cube_bytes == cube_bytes && hp.work_bytes >= work_need) return RML_OK;
  cudaDeviceSynchronize();
  free_pipe(hp);
  hp.work_bytes = work_need;
  for (int i = 0; i < kHostBufs; ++i) {
    RML_CUDA(c, cudaMalloc(&hp.cubes[i], chunk * cube_bytes));
    RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.ijk[i]), chunk * 3 * 4));
    RML_CUDA(c, cudaMalloc(&hp.work[i], hp.work_bytes));
    RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.proba[i]), chunk * kMaxClasses * 4));
    RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.label[i]), chunk * 4));
    RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.known[i]), chunk));
    RML_CUDA(c, cudaStreamCreateWithFlags(&hp.stream[i], cudaStreamNonBlocking));
    RML_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&hp.res_pin[i]), chunk * (kMaxClasses * 4 + 4 + 1), cudaHostAllocDefault));
    RML_CUDA(c, cudaEventCreateWithFlags(&hp.res_ev[i], cudaEventDisableTiming));
  }
  hp.chunk = chunk;
  hp.cube_bytes = cube_bytes;
  if (!cube_u8 && c->host_narrow) {
    // float32 cubes of the sensor's integers cross the bus as bytes (see host_narrow.h)
    for (int i = 0; i < kHostBufs; ++i) {
      RML_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&hp.narrow_pin[i]), chunk * (cube_bytes / 4), cudaHostAllocDefault));
      RML_CUDA(c, cudaEventCreateWithFlags(&hp.narrow_ev[i], cudaEventDisableTiming));
    }
    hp.pool = rml_host::narrow_pool_create(c->host_narrow_threads);
  }
  return RML_OK;
}

// workspace of the small calls: [u8 rows][int32 norms][digit planes + norms], all three present so
// that the scorer can fall through u8 -> digits -> float64 without another allocation
size_t small_u8_bytes(const rml_ctx* c, int64_t rows) {
  const int F = c->model.kind ? c->model.F : feature_len(c, RML_MASK_ALL);
  return align256(static_cast<size_t>(rows) * round_up(F, 128)) + align256(static_cast<size_t>(rows) * 4);
}
size_t small_work_bytes(const rml_ctx* c, int64_t rows) {
  return small_u8_bytes(c, rows) + digit_scratch_bytes(c, rows) + 256;
}

int reserve_small(rml_ctx* c, int64_t rows) {
  SmallPipe& sp = c->small;
  const int F = c->model.kind ? c->model.F : feature_len(c, RML_MASK_ALL);
  if (rows > kSmallMax) rows = kSmallMax;
  if (rows < 1) rows = 1;
  const size_t work_need = small_work_bytes(c, rows);
  if (sp.cap >= rows && sp.F == F && sp.work_bytes >= work_need) return RML_OK;
  cudaDeviceSynchronize();
  free_small(sp);
  RML_CUDA(c, cudaMalloc(&sp.cube, static_cast<size_t>(c->sx) * c->sy * c->sz * 4));
  RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&sp.ijk), rows * 12));
  RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&sp.feats), static_cast<size_t>(rows) * F * 4));
  RML_CUDA(c, cudaMalloc(&sp.work, work_need));
  RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&sp.proba), rows * kMaxClasses * 4));
  RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&sp.label), rows * 4));
  RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&sp.known), rows));
  sp.pinned_bytes = static_cast<size_t>(rows) * (kMaxClasses * 4 + 4 + 1) + 64;
  RML_CUDA(c, cudaHostAlloc(reinterpret_cast<void**>(&sp.pinned), sp.pinned_bytes, cudaHostAllocDefault));
  sp.cap = rows; sp.F = F; sp.work_bytes = work_need;
  return RML_OK;
}

// Host buffers in, host buffers out: the batch goes through the GPU in chunks on kHostBufs
// streams so that the H2D copy of chunk n+1 overlaps the kernels and the D2H copy of chunk n.
int predict_host_impl(rml_ctx* c, const void* cubes_host, int cube_u8, int64_t B, int mode,
                      const int32_t* ijk_host, uint32_t mask, double min_proba, float* proba_host,
                      int32_t* label_host, uint8_t* known_host) {
  if (!c) return RML_E_INVALID;
  if (c->model.kind == 0) return fail(c, RML_E_NOMODEL, "rml_predict_host: no model loaded");
  if (!cubes_host || !proba_host || !label_host || B < 0)
    return fail(c, RML_E_INVALID, "rml_predict_host: null buffer or B<0");
  if (mode == RML_MODE_SLICE && !ijk_host) return fail(c, RML_E_INVALID, "SLICE mode needs ijk");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  HostPipe& hp = c->pipe;
  const int64_t chunk = cube_u8 ? 1024 : 512;
  const size_t cube_bytes = static_cast<size_t>(c->sx) * c->sy * c->sz * (cube_u8 ? 1 : 4);
  const int C = c->model.C;
  if (hp.chunk != chunk || hp.cube_bytes != cube_bytes || hp.work_bytes < rml_predict_workspace_bytes(c, chunk))
    return fail(c, RML_E_INVALID, "rml_predict_host: staging buffers not reserved for this model / cube type; call "
                "rml_reserve(ctx, 0, %s) after loading the model", cube_u8 ? "RML_RESERVE_HOST_U8" : "RML_RESERVE_HOST");
  const char* src = static_cast<const char*>(cubes_host);
  int64_t done = 0;
  int slot = 0;
  // integral float32 cubes go over the bus as bytes when the model is on the integer path: the
  // conversion of chunk n+1 (host threads) overlaps the copy and the kernels of chunk n
  const bool narrow = !cube_u8 && hp.pool && !hp.narrow_off && use_u8_path(c);
  const size_t cube_elems = cube_bytes / (cube_u8 ? 1 : 4);
  hp.last_h2d_bytes = 0;
  hp.last_narrowed = 0;
  // results of the chunk that used a slot before: pinned mirror -> the caller's arrays
  auto drain = [&](int sl) -> int {
    if (!hp.res_n[sl]) return RML_OK;
    RML_CUDA(c, cudaEventSynchronize(hp.res_ev[sl]));
    const int64_t at = hp.res_at[sl], m = hp.res_n[sl];
    const unsigned char* r = hp.res_pin[sl];
    memcpy(proba_host + at * C, r, static_cast<size_t>(m) * C * 4);
    memcpy(label_host + at, r + chunk * kMaxClasses * 4, static_cast<size_t>(m) * 4);
    if (known_host) memcpy(known_host + at, r + chunk * (kMaxClasses * 4 + 4), static_cast<size_t>(m));
    hp.res_n[sl] = 0;
    return RML_OK;
  };
  for (int i = 0; i < kHostBufs; ++i) hp.res_n[i] = 0;
  while (done < B) {
    const int64_t n = (B - done) < chunk ? (B - done) : chunk;
    cudaStream_t st = hp.stream[slot];
    int as_u8 = cube_u8;
    { int rc = drain(slot); if (rc) return rc; }
    if (narrow && !hp.narrow_off && n >= 32) {      // a few scans are quicker to copy than to hand to the pool
      const auto tw = std::chrono::steady_clock::now();
      RML_CUDA(c, cudaEventSynchronize(hp.narrow_ev[slot]));     // the previous copy out of this staging buffer
      const auto t0 = std::chrono::steady_clock::now();
      const int bad = rml_host::narrow_f32_to_u8(hp.pool, reinterpret_cast<const float*>(src + done * cube_bytes),
                                                 hp.narrow_pin[slot], static_cast<size_t>(n) * cube_elems);
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      hp.narrow_gbs = static_cast<double>(n) * cube_bytes / sec * 1e-9;
      static const bool trace = getenv("RML_HOST_NARROW_TRACE") != nullptr;
      if (trace)
        fprintf(stderr, "[narrow] chunk at %lld: waited %.3f ms for the staging buffer, converted %lld scans in %.3f ms (%.1f GB/s), %s\n",
                static_cast<long long>(done), std::chrono::duration<double>(t0 - tw).count() * 1e3,
                static_cast<long long>(n), sec * 1e3, hp.narrow_gbs, bad ? "NOT integral" : "integral");
      // narrowing only pays while the host converts faster than the bus would move the float32 bytes
      // (PCIe 5 x16: ~53 GB/s); a busy or narrow host switches it off until the next rml_reserve
      if (n == chunk) {
        hp.narrow_slow = hp.narrow_gbs < c->host_narrow_min_gbs ? hp.narrow_slow + 1 : 0;
        if (hp.narrow_slow >= 3) hp.narrow_off = true;
      }
      if (bad && ++hp.narrow_bad >= 2) hp.narrow_off = true;     // real-valued cubes: stop paying for the check
      if (!bad) {
        RML_CUDA(c, cudaMemcpyAsync(hp.cubes[slot], hp.narrow_pin[slot], static_cast<size_t>(n) * cube_elems,
                                    cudaMemcpyHostToDevice, st));
        RML_CUDA(c, cudaEventRecord(hp.narrow_ev[slot], st));
        as_u8 = 1;
        hp.last_h2d_bytes += n * static_cast<int64_t>(cube_elems);
        hp.last_narrowed += n;
      }
    }
    if (as_u8 == cube_u8) {
      RML_CUDA(c, cudaMemcpyAsync(hp.cubes[slot], src + done * cube_bytes, n * cube_bytes,
                                  cudaMemcpyHostToDevice, st));
      hp.last_h2d_bytes += n * static_cast<int64_t>(cube_bytes);
    }
    if (mode == RML_MODE_SLICE)
      RML_CUDA(c, cudaMemcpyAsync(hp.ijk[slot], ijk_host + done * 3, n * 12, cudaMemcpyHostToDevice, st));
    int rc = predict_impl(c, hp.cubes[slot], n, mode, hp.ijk[slot], mask, min_proba, hp.work[slot],
                          hp.proba[slot], hp.label[slot], hp.known[slot], st, as_u8);
    if (rc) return rc;
    unsigned char* r = hp.res_pin[slot];
    RML_CUDA(c, cudaMemcpyAsync(r, hp.proba[slot], n * C * 4, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaMemcpyAsync(r + chunk * kMaxClasses * 4, hp.label[slot], n * 4, cudaMemcpyDeviceToHost, st));
    if (known_host)
      RML_CUDA(c, cudaMemcpyAsync(r + chunk * (kMaxClasses * 4 + 4), hp.known[slot], n, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaEventRecord(hp.res_ev[slot], st));
    hp.res_at[slot] = done; hp.res_n[slot] = n;
    done += n;
    slot = (slot + 1) % kHostBufs;
  }
  for (int i = 0; i < kHostBufs; ++i) {
    int rc = drain(i);
    if (rc) return rc;
  }
  for (int i = 0; i < kHostBufs; ++i) RML_CUDA(c, cudaStreamSynchronize(hp.stream[i]));
  return RML_OK;
}

// status word -> error code (shared by rml_check_status and the host-buffer calls)
int status_to_rc(rml_ctx* c, const unsigned int h[4]) {
  if (h[3]) return fail(c, RML_E_INVALID, "derive_targets: %u axis rankings had no finite maximum (NaN / -inf sums); the lowest unused index was emitted", h[3]);
  if (h[1]) return fail(c, RML_E_INVALID, "SLICE mode: %u scans had a target index outside the cube (numpy IndexError)", h[1]);
  if (h[2]) return fail(c, RML_E_RANGE, "digit path: %u scans had features outside [0, 256/scale); score them as RML_F32_EXACT", h[2]);
  if (h[0]) return fail(c, RML_E_NONINTEGRAL, "u8 path: %u warps saw values that are not integers in [0,255] (NaN included); use RML_F32", h[0]);
  return RML_OK;
}

// B <= kSmallMax rows already on the device in sp.feats (float32, scaled): score them with the
// best exact path and bring results + status back with ONE synchronisation in the common case.
// Order tried: u8 tensor-core scorer (integral model and integral rows) -> multi-digit
// tensor-core scorer -> float64 CUDA-core scorer; the switch is made on the device status word,
// never silently on the host.
int score_small(rml_ctx* c, int64_t B, double min_proba, float* proba_host, int32_t* label_host,
                uint8_t* known_host, cudaStream_t st) {
  SmallPipe& sp = c->small;
  Model& m = c->model;
  const int C = m.C;
  unsigned char* pin = sp.pinned;
  float* pin_proba = reinterpret_cast<float*>(pin);
  int32_t* pin_label = reinterpret_cast<int32_t*>(pin + static_cast<size_t>(sp.cap) * kMaxClasses * 4);
  uint8_t* pin_known = reinterpret_cast<uint8_t*>(pin_label + sp.cap);
  unsigned int* pin_status = reinterpret_cast<unsigned int*>(pin + sp.pinned_bytes - 16);
  const bool can_u8 = m.kind == 1 && m.integral && !c->force_f32 && !c->aff_off_dev;
  char* ws = static_cast<char*>(sp.work);
  for (int attempt = can_u8 ? 0 : 1; attempt < 3; ++attempt) {
    int rc;
    if (attempt == 0) {
      uint8_t* q = reinterpret_cast<uint8_t*>(ws);
      int32_t* norms = reinterpret_cast<int32_t*>(ws + align256(static_cast<size_t>(sp.cap) * m.kpad));
      QuantParams qp;
      qp.feats = sp.feats; qp.out = q; qp.norms = norms; qp.status = c->status; qp.B = B; qp.F = m.F;
      qp.stride = m.kpad; qp.scale = static_cast<float>(m.feature_scale);
      k1_quantize<<<static_cast<unsigned>((B + 7) / 8), 256, 0, st>>>(qp);
      RML_CUDA(c, cudaGetLastError());
      ++c->launches;
      rc = score_impl(c, q, RML_U8, norms, B, min_proba, sp.proba, nullptr, sp.label, sp.known, st);
    } else if (attempt == 1) {
      if (m.kind == 1 && !m.digits_ok) continue;
      rc = score_impl(c, sp.feats, RML_F32, nullptr, B, min_proba, sp.proba, nullptr, sp.label, sp.known, st, 0,
                      nullptr, m.kind == 1 ? ws + small_u8_bytes(c, sp.cap) : nullptr);
    } else {
      rc = score_impl(c, sp.feats, RML_F32_EXACT, nullptr, B, min_proba, sp.proba, nullptr, sp.label, sp.known, st);
    }
    if (rc) return rc;
    RML_CUDA(c, cudaMemcpyAsync(pin_proba, sp.proba, B * C * 4, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaMemcpyAsync(pin_label, sp.label, B * 4, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaMemcpyAsync(pin_known, sp.known, B, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaMemcpyAsync(pin_status, c->status, 16, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaMemsetAsync(c->status, 0, 16, st));
    RML_CUDA(c, cudaStreamSynchronize(st));
    const bool retry = (attempt == 0 && pin_status[0]) || (attempt == 1 && pin_status[2]);
    if (retry && !pin_status[1] && !pin_status[3]) continue;
    int src = status_to_rc(c, pin_status);
    if (src) return src;
    memcpy(proba_host, pin_proba, B * C * 4);
    memcpy(label_host, pin_label, B * 4);
    if (known_host) memcpy(known_host, pin_known, B);
    return RML_OK;
  }
  return fail(c, RML_E_UNSUPPORTED, "score_small: no scorer accepted the features");
}
}  // namespace

extern "C" {

int rml_reserve(rml_ctx* c, int64_t max_batch, int flags) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  if (flags & RML_RESERVE_SCORE) {
    const Model& m = c->model;
    if (m.kind == 1 && m.digits_ok && c->dg_cap < max_batch) {
      cudaDeviceSynchronize();
      cudaFree(c->dg_planes); cudaFree(c->dg_norms);
      c->dg_planes = nullptr; c->dg_norms = nullptr; c->dg_cap = 0;
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->dg_planes), align256(static_cast<size_t>(3) * max_batch * m.kpad)));
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->dg_norms), static_cast<size_t>(max_batch) * 8));
      c->dg_cap = max_batch;
    }
  }
  int rc;
  if ((flags & RML_RESERVE_HOST) && (rc = reserve_host_pipe(c, 0))) return rc;
  if ((flags & RML_RESERVE_HOST_U8) && (rc = reserve_host_pipe(c, 1))) return rc;
  if ((flags & RML_RESERVE_SMALL) && (rc = reserve_small(c, max_batch))) return rc;
  return RML_OK;
}

int rml_predict_host(rml_ctx* c, const float* cubes_host, int64_t B, int mode,
                     const int32_t* ijk_host, uint32_t mask, double min_proba, float* proba_host,
                     int32_t* label_host, uint8_t* known_host) {
  return predict_host_impl(c, cubes_host, 0, B, mode, ijk_host, mask, min_proba, proba_host,
                           label_host, known_host);
}

int rml_set_host_narrowing(rml_ctx* c, int enabled, int threads, double min_gbs) {
  if (!c) return RML_E_INVALID;
  c->host_narrow = enabled ? 1 : 0;
  c->host_narrow_threads = threads > 0 ? threads : 0;
  if (min_gbs > 0) c->host_narrow_min_gbs = min_gbs;
  // takes effect at the next rml_reserve(RML_RESERVE_HOST): drop the current staging
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  free_pipe(c->pipe);
  return RML_OK;
}

int rml_host_narrow_f32_to_u8(const float* src_host, uint8_t* dst_host, int64_t n, int threads) {
  if (!src_host || !dst_host || n < 0 || (reinterpret_cast<uintptr_t>(dst_host) & 31)) return RML_E_INVALID;
  rml_host::NarrowPool* pool = threads == 1 ? nullptr : rml_host::narrow_pool_create(threads);
  const int bad = rml_host::narrow_f32_to_u8(pool, src_host, dst_host, static_cast<size_t>(n));
  rml_host::narrow_pool_destroy(pool);
  return bad ? RML_E_NONINTEGRAL : RML_OK;
}

int rml_last_host_transfer(const rml_ctx* c, int64_t* h2d_bytes, int64_t* narrowed_scans, double* convert_gbs,
                           int* threads, int* active) {
  if (!c) return RML_E_INVALID;
  const HostPipe& hp = c->pipe;
  if (h2d_bytes) *h2d_bytes = hp.last_h2d_bytes;
  if (narrowed_scans) *narrowed_scans = hp.last_narrowed;
  if (convert_gbs) *convert_gbs = hp.narrow_gbs;
  if (threads) *threads = rml_host::narrow_pool_threads(hp.pool);
  if (active) *active = hp.pool && !hp.narrow_off;
  return RML_OK;
}

int rml_predict_host_u8(rml_ctx* c, const uint8_t* cubes_host, int64_t B, int mode,
                        const int32_t* ijk_host, uint32_t mask, double min_proba, float* proba_host,
                        int32_t* label_host, uint8_t* known_host) {
  return predict_host_impl(c, cubes_host, 1, B, mode, ijk_host, mask, min_proba, proba_host,
                           label_host, known_host);
}

// predict.py:56-70 for host features: model.predict_proba(X) + argmax + threshold
int rml_score_host(rml_ctx* c, const float* feats_host, int64_t B, double min_proba, float* proba_host,
                   int32_t* label_host, uint8_t* known_host) {
  if (!c) return RML_E_INVALID;
  if (c->model.kind == 0) return fail(c, RML_E_NOMODEL, "rml_score_host: no model loaded");
  if (!feats_host || !proba_host || !label_host || B < 0) return fail(c, RML_E_INVALID, "rml_score_host: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  SmallPipe& sp = c->small;
  if (sp.cap < 1 || sp.F != c->model.F || sp.work_bytes < small_work_bytes(c, sp.cap))
    return fail(c, RML_E_INVALID, "rml_score_host: call rml_reserve(ctx, rows, RML_RESERVE_SMALL) after loading the model");
  cudaStream_t st = c->aux_stream;
  const int F = c->model.F, C = c->model.C;
  for (int64_t done = 0; done < B; done += sp.cap) {
    const int64_t n = (B - done) < sp.cap ? (B - done) : sp.cap;
    RML_CUDA(c, cudaMemcpyAsync(sp.feats, feats_host + done * F, static_cast<size_t>(n) * F * 4, cudaMemcpyHostToDevice, st));
    int rc = score_small(c, n, min_proba, proba_host + done * C, label_host + done,
                         known_host ? known_host + done : nullptr, st);
    if (rc) return rc;
  }
  return RML_OK;
}

// One predict.py:93-119 iteration: ONE raw cube (predict.py:90-91) and its T detected targets.
// The cube crosses PCIe once; every target takes its three slices from it (predict.py:102-107),
// concat + /255 (common.py:141-149) and the classifier chain, all behind one synchronisation.
int rml_predict_targets_host(rml_ctx* c, const float* cube_host, int T, const int32_t* ijk_host,
                             uint32_t mask, double min_proba, float* proba_host, int32_t* label_host,
                             uint8_t* known_host) {
  if (!c) return RML_E_INVALID;
  if (c->model.kind == 0) return fail(c, RML_E_NOMODEL, "rml_predict_targets_host: no model loaded");
  if (!cube_host || !ijk_host || !proba_host || !label_host || T < 0)
    return fail(c, RML_E_INVALID, "rml_predict_targets_host: null buffer or T<0");
  if (T == 0) return RML_OK;
  if (feature_len(c, mask) != c->model.F)
    return fail(c, RML_E_INVALID, "rml_predict_targets_host: mask gives F=%d but the model has F=%d", feature_len(c, mask), c->model.F);
  DeviceGuard g(c->device);
  SmallPipe& sp = c->small;
  if (sp.cap < T || sp.F != c->model.F || sp.work_bytes < small_work_bytes(c, sp.cap))
    return fail(c, RML_E_INVALID, "rml_predict_targets_host: call rml_reserve(ctx, max_targets, RML_RESERVE_SMALL) after loading the model");
  cudaStream_t st = c->aux_stream;
  RML_CUDA(c, cudaMemcpyAsync(sp.cube, cube_host, static_cast<size_t>(c->sx) * c->sy * c->sz * 4, cudaMemcpyHostToDevice, st));
  RML_CUDA(c, cudaMemcpyAsync(sp.ijk, ijk_host, static_cast<size_t>(T) * 12, cudaMemcpyHostToDevice, st));
  // float32 feature rows exactly as common.process_samples(scale=True) returns them, then the
  // scorer chain picks the fastest exact path on the device status
  const Affine aff = c->aff_off_dev ? Affine{0.f, 1.f, 1, c->aff_off_dev, c->aff_scl_dev}
                                    : Affine{0.f, static_cast<float>(c->model.feature_scale), 1};
  int rc = project_impl(c, sp.cube, T, RML_MODE_SLICE, sp.ijk, mask, RML_F32, sp.feats, nullptr, st, 0, nullptr,
                        &aff, 0, /*cube_stride=*/0);
  if (rc) return rc;
  return score_small(c, T, min_proba, proba_host, label_host, known_host, st);
}

// ------------------------------------------------------------------------------ label exchange
static int nccl_load(rml_ctx* c) {
  if (g_nccl.handle) return RML_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (torch)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    if (const char* e = getenv("RML_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(c, RML_E_UNSUPPORTED, "NCCL not found (libnccl.so.2 / RML_NCCL_LIB): %s", dlerror());
  g_nccl.GetUniqueId = reinterpret_cast<int (*)(rml_nccl_uid*)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, rml_nccl_uid, int)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(h, "ncclAllGather"));
  g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy)
    return fail(c, RML_E_UNSUPPORTED, "NCCL symbols missing in the loaded library");
  g_nccl.handle = h;
  return RML_OK;
}

int rml_comm_unique_id(rml_ctx* c, void* id128) {
  if (!c || !id128) return RML_E_INVALID;
  int rc = nccl_load(c);
  if (rc) return rc;
  rml_nccl_uid id;
  const int r = g_nccl.GetUniqueId(&id);
  if (r) return fail(c, RML_E_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  memcpy(id128, &id, 128);
  return RML_OK;
}

int rml_comm_init(rml_ctx* c, int rank, int world, const void* id128) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(c, RML_E_INVALID, "rml_comm_init: bad arguments");
  int rc = nccl_load(c);
  if (rc) return rc;
  DeviceGuard g(c->device);
  if (c->nccl_comm) { g_nccl.CommDestroy(c->nccl_comm); c->nccl_comm = nullptr; }
  rml_nccl_uid id;
  memcpy(&id, id128, 128);
  const int r = g_nccl.CommInitRank(&c->nccl_comm, world, id, rank);
  if (r) return fail(c, RML_E_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  c->nccl_rank = rank; c->nccl_world = world;
  return RML_OK;
}

int rml_comm_destroy(rml_ctx* c) {
  if (!c) return RML_E_INVALID;
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
  c->nccl_comm = nullptr; c->nccl_world = 1; c->nccl_rank = 0;
  return RML_OK;
}

// ONE all-gather of the int32 labels per batch.  send_dev may be the rank's own
// slice of recv_dev (recv_dev + rank * count): rml_predict then writes the labels straight into
// the gather buffer and NCCL runs in place, no copy between the scorer and the collective.
int rml_allgather_labels(rml_ctx* c, const int32_t* send_dev, int32_t* recv_dev, int64_t count, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!send_dev || !recv_dev || count < 0) return fail(c, RML_E_INVALID, "rml_allgather_labels: null buffer or count<0");
  DeviceGuard g(c->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->nccl_world == 1 && !c->nccl_comm) {
    if (send_dev != recv_dev)
      RML_CUDA(c, cudaMemcpyAsync(recv_dev, send_dev, static_cast<size_t>(count) * 4, cudaMemcpyDeviceToDevice, st));
    return RML_OK;
  }
  if (!c->nccl_comm) return fail(c, RML_E_INVALID, "rml_allgather_labels: call rml_comm_init first");
  const int r = g_nccl.AllGather(send_dev, recv_dev, static_cast<size_t>(count), /*ncclInt32*/ 2, c->nccl_comm, st);
  if (r) return fail(c, RML_E_CUDA, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return RML_OK;
}

int rml_check_status(rml_ctx* c, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  unsigned int h[4] = {0, 0, 0, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RML_CUDA(c, cudaMemcpyAsync(h, c->status, 16, cudaMemcpyDeviceToHost, st));
  RML_CUDA(c, cudaMemsetAsync(c->status, 0, 16, st));
  RML_CUDA(c, cudaStreamSynchronize(st));
  return status_to_rc(c, h);
}

int64_t rml_launch_count(const rml_ctx* c) { return c ? c->launches : 0; }

int rml_enable_timing(rml_ctx* c, int enabled) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  if (enabled && !c->ev_k1a) {
    RML_CUDA(c, cudaEventCreate(&c->ev_k1a));
    RML_CUDA(c, cudaEventCreate(&c->ev_k1b));
    RML_CUDA(c, cudaEventCreate(&c->ev_k2b));
  } else if (!enabled && c->ev_k1a) {
    cudaEventDestroy(c->ev_k1a); cudaEventDestroy(c->ev_k1b); cudaEventDestroy(c->ev_k2b);
    c->ev_k1a = c->ev_k1b = c->ev_k2b = nullptr;
  }
  return RML_OK;
}

int rml_last_timing(rml_ctx* c, float* k1_ms, float* total_ms, int* fused) {
  if (!c) return RML_E_INVALID;
  if (!c->ev_k1a) return fail(c, RML_E_INVALID, "rml_last_timing: call rml_enable_timing first");
  DeviceGuard g(c->device);
  RML_CUDA(c, cudaEventSynchronize(c->ev_k2b));
  if (k1_ms) RML_CUDA(c, cudaEventElapsedTime(k1_ms, c->ev_k1a, c->ev_k1b));
  if (total_ms) RML_CUDA(c, cudaEventElapsedTime(total_ms, c->ev_k1a, c->ev_k2b));
  if (fused) *fused = c->last_fused;
  return RML_OK;
}

int rml_set_fused(rml_ctx* c, int enabled, int k2_sms, int64_t min_batch) {
  if (!c) return RML_E_INVALID;
  c->fused_enabled = enabled;
  if (k2_sms > 0) c->k2_sms = k2_sms;
  if (min_batch > 0) c->fused_min_b = min_batch;
  return RML_OK;
}

int rml_set_fused_u8(rml_ctx* c, int k2_sms) {
  if (!c) return RML_E_INVALID;
  if (k2_sms < 0 || k2_sms >= c->num_sms) return fail(c, RML_E_INVALID, "rml_set_fused_u8: k2_sms %d not in [0, %d)", k2_sms, c->num_sms);
  c->k2_sms_u8 = k2_sms;
  return RML_OK;
}


// ------------------------------------------------------------------------------ §8f callers
int rml_derive_targets(rml_ctx* c, const float* cubes, int64_t B, int num_targets, int32_t* ijk,
                       float* sums, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!cubes || !ijk || B < 0 || num_targets < 1 || num_targets > kMaxTargets ||
      num_targets > c->sx || num_targets > c->sy || num_targets > c->sz)
    return fail(c, RML_E_INVALID, "rml_derive_targets: bad arguments");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  DeriveParams p;
  p.cubes = cubes; p.ijk = ijk; p.sums = sums; p.status = c->status; p.B = B; p.sx = c->sx; p.sy = c->sy; p.sz = c->sz;
  p.T = num_targets;
  const int smem = (c->sx + c->sy + c->sz + 8 * c->sz + c->sx * c->sy) * 4;
  const int grid = static_cast<int>(B < 8ll * c->num_sms ? B : 8ll * c->num_sms);
  k0_derive_targets<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

int rml_project_derive(rml_ctx* c, const float* cubes, int64_t B, uint32_t mask, int dtype, void* feats,
                       int32_t* norms, int num_targets, int32_t* ijk, float* sums, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!ijk || num_targets < 1 || num_targets > kK1MaxTargets || num_targets > c->sx || num_targets > c->sy ||
      num_targets > c->sz)
    return fail(c, RML_E_INVALID, "rml_project_derive: bad arguments");
  DeviceGuard g(c->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->sx == kSX && c->sy == kSY && c->sz == kSZ) {
    // ONE pass over the cube: the projection kernel's row warps also accumulate the three axis sums
    const DeriveOut dv{num_targets, ijk, sums};
    return project_impl(c, cubes, B, RML_MODE_MAX, nullptr, mask, dtype, feats, norms, st, 0, nullptr, nullptr, 0, -1, &dv);
  }
  // other arenas: the generic projection kernel, then the stand-alone axis-sum kernel (two passes)
  int rc = project_impl(c, cubes, B, RML_MODE_MAX, nullptr, mask, dtype, feats, norms, st);
  if (rc) return rc;
  return rml_derive_targets(c, cubes, B, num_targets, ijk, sums, stream);
}

int rml_set_zoom(rml_ctx* c, int proj, int in_h, int in_w, int out_h, int out_w,
                 const double* a_rows, const double* a_cols) {
  if (!c) return RML_E_INVALID;
  if (proj < 0 || proj > 2 || in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0 || !a_rows || !a_cols)
    return fail(c, RML_E_INVALID, "rml_set_zoom: bad arguments");
  if (static_cast<size_t>(in_h) * out_w * 8 + static_cast<size_t>(in_h) * in_w * 4 > 200 * 1024)
    return fail(c, RML_E_UNSUPPORTED, "rml_set_zoom: projection too large for the smem-resident zoom");
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  cudaFree(c->zoom_ar[proj]); cudaFree(c->zoom_ac[proj]);
  c->zoom_ar[proj] = c->zoom_ac[proj] = nullptr;
  int rc;
  if ((rc = upload(c, &c->zoom_ar[proj], a_rows, static_cast<size_t>(out_h) * in_h))) return rc;
  if ((rc = upload(c, &c->zoom_ac[proj], a_cols, static_cast<size_t>(out_w) * in_w))) return rc;
  c->zoom_ih[proj] = in_h; c->zoom_iw[proj] = in_w; c->zoom_oh[proj] = out_h; c->zoom_ow[proj] = out_w;
  return RML_OK;
}

int rml_zoom_feature_len(const rml_ctx* c, uint32_t mask) {
  if (!c) return RML_E_INVALID;
  int f = 0;
  for (int q = 0; q < 3; ++q)
    if (mask & (1u << q)) f += c->zoom_oh[q] * c->zoom_ow[q];
  return f;
}

int rml_process_samples_zoom(rml_ctx* c, const float* xz, int64_t stride_xz, const float* yz,
                             int64_t stride_yz, const float* xy, int64_t stride_xy, int64_t B,
                             uint32_t mask, int scale, float* feats, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!feats || B < 0 || (mask & RML_MASK_ALL) == 0)
    return fail(c, RML_E_INVALID, "rml_process_samples_zoom: bad arguments");
  const float* src[3] = {xz, yz, xy};
  const int64_t strides[3] = {stride_xz, stride_yz, stride_xy};
  for (int q = 0; q < 3; ++q)
    if ((mask & (1u << q)) && (!src[q] || !c->zoom_ar[q]))
      return fail(c, RML_E_INVALID, "rml_process_samples_zoom: projection %d missing or rml_set_zoom not called", q);
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  ZoomParams p;
  int off = 0, smem = 0;
  for (int q = 0; q < 3; ++q) {
    const bool on = mask & (1u << q);
    p.proj[q] = on ? src[q] : nullptr;
    p.pstride[q] = strides[q];
    p.ih[q] = c->zoom_ih[q]; p.iw[q] = c->zoom_iw[q]; p.oh[q] = c->zoom_oh[q]; p.ow[q] = c->zoom_ow[q];
    p.ar[q] = c->zoom_ar[q]; p.ac[q] = c->zoom_ac[q];
    p.off[q] = off;
    if (on) {
      off += p.oh[q] * p.ow[q];
      const int sm = p.ih[q] * p.ow[q] * 8 + p.ih[q] * p.iw[q] * 4;
      if (sm > smem) smem = sm;
    }
  }
  p.feats = feats; p.B = B; p.F = off; p.scale = scale;
  p.offset = c->aff_offset; p.scale_value = c->aff_scale;
  p.aff_off = c->aff_off_dev; p.aff_scl = c->aff_scl_dev;
  if (scale && c->aff_off_dev && c->aff_F != p.F)
    return fail(c, RML_E_INVALID, "rml_process_samples_zoom: the per-feature affine has F=%d, the zoomed row F=%d", c->aff_F, p.F);
  const int gx = static_cast<int>(B < 4ll * c->num_sms ? B : 4ll * c->num_sms);
  k0_zoom_concat<<<dim3(gx, 3), 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

// ------------------------------------------------------------------------------ networks
int rml_net_begin(rml_ctx* c, int resize_to, int n_classes, int head, float alpha) {
  if (!c) return RML_E_INVALID;
  if (resize_to < 8 || resize_to > 256 || n_classes < 1 || n_classes > 8 || head < 0 || head > 1)
    return fail(c, RML_E_INVALID, "rml_net_begin: bad arguments");
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  free_net(c->net);
  c->net.R = resize_to; c->net.C = n_classes; c->net.head = head; c->net.alpha = alpha;
  return RML_OK;
}

int rml_net_set_resize_tables(rml_ctx* c, int branch, int ksh, const double* kh, const int32_t* bh,
                              int ksv, const double* kv, const int32_t* bv) {
  if (!c) return RML_E_INVALID;
  Net& n = c->net;
  if (branch < 0 || branch > 2 || n.R == 0 || !kh || !bh || !kv || !bv || ksh <= 0 || ksv <= 0)
    return fail(c, RML_E_INVALID, "rml_net_set_resize_tables: bad arguments");
  DeviceGuard g(c->device);
  int rc;
  if ((rc = upload(c, &n.kh[branch], kh, static_cast<size_t>(n.R) * ksh))) return rc;
  if ((rc = upload(c, &n.kv[branch], kv, static_cast<size_t>(n.R) * ksv))) return rc;
  if ((rc = upload(c, reinterpret_cast<int32_t**>(&n.bh[branch]), bh, static_cast<size_t>(n.R) * 2))) return rc;
  if ((rc = upload(c, reinterpret_cast<int32_t**>(&n.bv[branch]), bv, static_cast<size_t>(n.R) * 2))) return rc;
  n.ksh[branch] = ksh; n.ksv[branch] = ksv;
  return RML_OK;
}

int rml_net_add_conv(rml_ctx* c, int layer, int branch, int cin, int cout, int act,
                     const float* w_hwio, const float* bias) {
  if (!c) return RML_E_INVALID;
  Net& n = c->net;
  if (layer < 0 || layer > 8 || branch < 0 || branch > 2 || cin <= 0 || cout <= 0 ||
      cout % kConvCoutTile || act < 0 || act > 2 || !w_hwio || !bias)
    return fail(c, RML_E_INVALID, "rml_net_add_conv: bad arguments (cout must be a multiple of 32)");
  if (9 * cin * kConvCoutTile * 4 > 200 * 1024)
    return fail(c, RML_E_UNSUPPORTED, "rml_net_add_conv: cin=%d too large for the smem-resident weights", cin);
  if (static_cast<int>(n.convs.size()) <= layer) n.convs.resize(layer + 1);
  NetConv& cv = n.convs[layer];
  if (cv.cin && (cv.cin != cin || cv.cout != cout || cv.act != act))
    return fail(c, RML_E_INVALID, "rml_net_add_conv: branches of layer %d disagree", layer);
  cv.cin = cin; cv.cout = cout; cv.act = act;
  DeviceGuard g(c->device);
  int rc;
  if ((rc = upload(c, &cv.w[branch], w_hwio, static_cast<size_t>(9) * cin * cout))) return rc;
  if ((rc = upload(c, &cv.bias[branch], bias, static_cast<size_t>(cout)))) return rc;
  cv.w_host[branch].assign(w_hwio, w_hwio + static_cast<size_t>(9) * cin * cout);
  cv.b_host[branch].assign(bias, bias + cout);
  return RML_OK;
}

int rml_net_set_dense(rml_ctx* c, int K, const uint16_t* w1t_bf16, const float* b1, int act1,
                      const float* w2, const float* b2, int act2, const float* w3, const float* b3) {
  if (!c) return RML_E_INVALID;
  Net& n = c->net;
  if (K <= 0 || K % kK5BlockK || !w1t_bf16 || !b1 || !w2 || !b2 || !w3 || !b3)
    return fail(c, RML_E_INVALID, "rml_net_set_dense: bad arguments (K must be a multiple of 64)");
  DeviceGuard g(c->device);
  int rc;
  if ((rc = upload(c, &n.w1t, w1t_bf16, static_cast<size_t>(64) * K))) return rc;
  if ((rc = upload(c, &n.b1, b1, 64))) return rc;
  if ((rc = upload(c, &n.w2, w2, 64 * 64))) return rc;
  if ((rc = upload(c, &n.b2, b2, 64))) return rc;
  if ((rc = upload(c, &n.w3, w3, static_cast<size_t>(64) * n.C))) return rc;
  if ((rc = upload(c, &n.b3, b3, n.C))) return rc;
  n.K = K; n.act1 = act1; n.act2 = act2;
  return RML_OK;
}

static int encode_bf16_map(rml_ctx* c, CUtensorMap* map, const void* base, int64_t rows, int K, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kK5BlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = c->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, RML_E_CUDA, "cuTensorMapEncodeTiled(bf16) failed: %d", (int)r);
  return RML_OK;
}

// spatial size after the conv tower and consistency of the declared shapes
static int net_out_hw(const Net& n) {
  int hw = n.R;
  for (size_t l = 0; l < n.convs.size(); ++l) hw = (hw + 1) / 2;
  return hw;
}

int rml_net_finish(rml_ctx* c) {
  if (!c) return RML_E_INVALID;
  Net& n = c->net;
  if (n.convs.empty() || !n.w1t) return fail(c, RML_E_INVALID, "rml_net_finish: layers missing");
  int cin = 1;
  for (size_t l = 0; l < n.convs.size(); ++l) {
    const NetConv& cv = n.convs[l];
    for (int b = 0; b < 3; ++b)
      if (!cv.w[b]) return fail(c, RML_E_INVALID, "conv layer %zu branch %d missing", l, b);
    if (cv.cin != cin) return fail(c, RML_E_INVALID, "conv layer %zu: cin=%d, expected %d", l, cv.cin, cin);
    cin = cv.cout;
  }
  for (int b = 0; b < 3; ++b)
    if (!n.kh[b]) return fail(c, RML_E_INVALID, "resize tables of branch %d missing", b);
  const int hw = net_out_hw(n);
  if (3 * hw * hw * cin != n.K)
    return fail(c, RML_E_INVALID, "dense K=%d but the conv towers produce 3*%d*%d*%d", n.K, hw, hw, cin);
  DeviceGuard g(c->device);
  int rc = encode_bf16_map(c, &n.map_w1, n.w1t, 64, n.K, 64);
  if (rc) return rc;
  // tensor-core implicit GEMM for every layer after the first when the shapes allow it
  bool ok = n.convs.size() >= 2;
  {
    int hw = n.R;
    for (size_t l = 0; l < n.convs.size(); ++l) {
      const NetConv& cv = n.convs[l];
      const int ho = (hw + 1) / 2;
      if (l >= 1 && (cv.cin % 64 || cv.cout % 16 || cv.cout > 128 || ho > 128 || (hw & 1))) ok = false;
      hw = ho;
    }
  }
  if (const char* e = getenv("RML_IGEMM")) ok = ok && atoi(e) != 0;
  n.use_igemm = ok;
  if (const char* e = getenv("RML_NET_FUSE1")) n.fuse_resize = atoi(e) != 0;
  if (ok) {
    for (size_t l = 1; l < n.convs.size(); ++l) {
      NetConv& cv = n.convs[l];
      const int K = 9 * cv.cin;
      std::vector<uint16_t> wt(static_cast<size_t>(3) * cv.cout * K);
      std::vector<float> b3(static_cast<size_t>(3) * cv.cout);
      for (int br = 0; br < 3; ++br) {
        for (int co = 0; co < cv.cout; ++co) {
          b3[br * cv.cout + co] = cv.b_host[br][co];
          for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < cv.cin; ++ci) {
              const float v = cv.w_host[br][(static_cast<size_t>(tap) * cv.cin + ci) * cv.cout + co];
              uint32_t u;
              memcpy(&u, &v, 4);
              u = (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;      // round to nearest even
              wt[(static_cast<size_t>(br) * cv.cout + co) * K + tap * cv.cin + ci] = static_cast<uint16_t>(u);
            }
        }
      }
      if ((rc = upload(c, &cv.wt_bf16, wt.data(), wt.size()))) return rc;
      if ((rc = upload(c, &cv.bias3, b3.data(), b3.size()))) return rc;
      if ((rc = encode_bf16_map(c, &cv.map_w, cv.wt_bf16, 3 * cv.cout, K, cv.cout))) return rc;
    }
  }
  // fused tower kernel: dnn.py's towers (80x80 -> Conv 64 -> Conv 32) run entirely on chip; sgan.py's
  // first layer (128x128 -> Conv 128) moves to the tensor cores
  n.tower_mode = 0;
  if (ok && n.fuse_resize) {
    const NetConv& c0 = n.convs[0];
    if (n.R == 80 && n.convs.size() == 2 && c0.cout == 64 && n.convs[1].cout == 32 && c0.act == 1 &&
        n.convs[1].act == 1 && c->sx * c->sz <= 31 * 176 && c->sy * c->sz <= 31 * 176 && c->sx <= 31 && c->sy <= 31)
      n.tower_mode = 1;
    else if (n.R == 128 && c0.cout == 128 && c0.act == 2 && n.alpha > 0.f && n.alpha < 1.f && c->sx * c->sz <= 31 * 176 && c->sy * c->sz <= 31 * 176 &&
             c->sx <= 31 && c->sy <= 31)
      n.tower_mode = 2;
    if (const char* e = getenv("RML_NET_TOWER")) { if (atoi(e) == 0) n.tower_mode = 0; }
  }
  if (n.tower_mode) {
    const NetConv& c0 = n.convs[0];
    const int C1 = c0.cout;
    // fp16 hi / lo split (22 bits; fp16 subnormals bound the absolute error by 2^-25)
    auto f16_bits = [](float v) {
      const __half_raw r = static_cast<__half_raw>(__float2half_rn(v));
      return static_cast<uint16_t>(r.x);
    };
    auto f16_to_f = [](uint16_t b) {
      __half_raw r;
      r.x = b;
      return __half2float(__half(r));
    };
    std::vector<uint16_t> w1(static_cast<size_t>(3) * C1 * kT6K1, 0);
    std::vector<float> b1(static_cast<size_t>(3) * C1);
    for (int br = 0; br < 3; ++br)
      for (int co = 0; co < C1; ++co) {
        b1[br * C1 + co] = c0.b_host[br][co];
        uint16_t* row = &w1[(static_cast<size_t>(br) * C1 + co) * kT6K1];
        for (int tap = 0; tap < 9; ++tap) {
          const float w = c0.w_host[br][static_cast<size_t>(tap) * C1 + co];     // HWIO with cin = 1
          const uint16_t hi = f16_bits(w);
          const uint16_t lo = f16_bits(w - f16_to_f(hi));
          row[tap] = hi; row[10 + tap] = hi; row[20 + tap] = lo;                 // pairs with A: hi | lo | hi
        }
        // the bias rides in the GEMM: A carries 1.0 in columns 9 and 29
        const uint16_t bhi = f16_bits(c0.b_host[br][co]);
        row[9] = bhi;
        row[29] = f16_bits(c0.b_host[br][co] - f16_to_f(bhi));
      }
    if ((rc = upload(c, &n.t6_w1, w1.data(), w1.size()))) return rc;
    if ((rc = upload(c, &n.t6_b1, b1.data(), b1.size()))) return rc;
    // persistent CTAs per branch, proportional to the per-image cost (resize rows differ: 22 / 31 / 22)
    if (const char* e = getenv("RML_T6_DBG")) n.t6_dbg = atoi(e);
    if (const char* e = getenv("RML_K4_SHARE")) n.k4_share = atoi(e) != 0;
    int split[3] = {50, 52, 46};
    if (const char* e = getenv("RML_T6_SPLIT")) sscanf(e, "%d,%d,%d", &split[0], &split[1], &split[2]);
    const int tot = split[0] + split[1] + split[2];
    n.t6_ctas[0] = c->num_sms * split[0] / tot;
    n.t6_ctas[1] = c->num_sms * split[1] / tot;
    n.t6_ctas[2] = c->num_sms - n.t6_ctas[0] - n.t6_ctas[1];
    for (int b = 0; b < 3; ++b) if (n.t6_ctas[b] < 1) n.tower_mode = 0;
  }
  n.ready = true;
  return RML_OK;
}

// Workspace plan of the network forward:
//   [ tower chunk: images | ping | pong ] [ flat bf16 tower output of a dense group ] [ f32 features of a chunk ]
// The conv towers run chunk by chunk (their activations are large); the tcgen05 dense stack runs
// once per dense group of up to kDenseGroup scans so that it has one tile for every SM.
constexpr int64_t kDenseGroup = 148 * 128;   // one 128-scan tile per SM: the dense stack fills the chip
// per-scan bytes of the tower scratch: resized images (only the unfused resize kernel writes them)
// and the two ping-pong activation buffers (bf16 on the tensor-core path; none at all when the
// dnn towers run fused on chip)
static void net_tower_sizes(const Net& n, size_t* img, size_t* act) {
  *img = n.tower_mode ? 0 : static_cast<size_t>(3) * n.R * n.R * 4;
  *act = 0;
  if (n.tower_mode == 1) return;
  const size_t esz = n.use_igemm ? 2 : 4;
  int hw = n.R;
  for (size_t l = 0; l + 1 < n.convs.size(); ++l) {
    hw = (hw + 1) / 2;
    const size_t a = static_cast<size_t>(3) * hw * hw * n.convs[l].cout * esz;
    if (a > *act) *act = a;
  }
}
static size_t net_tower_bytes_per_scan(const rml_ctx* c) {
  size_t img, act;
  net_tower_sizes(c->net, &img, &act);
  return img + 2 * act + 64;
}
struct NetPlan {
  int64_t chunk = 0, group = 0;
  size_t flat_off = 0, feats_off = 0;
};
static size_t net_plan_bytes(const rml_ctx* c, int64_t chunk, int64_t group) {
  const size_t F4 = static_cast<size_t>(feature_len(c, RML_MASK_ALL)) * 4;
  return align256(net_tower_bytes_per_scan(c) * chunk + 1024) + align256(static_cast<size_t>(group) * c->net.K * 2) +
         align256(F4 * chunk) + 1024;
}
static int net_make_plan(rml_ctx* c, int64_t B, size_t ws_bytes, NetPlan* pl) {
  int64_t group = B < kDenseGroup ? B : kDenseGroup;
  for (;;) {
    const size_t flat = align256(static_cast<size_t>(group) * c->net.K * 2);
    const size_t F4 = static_cast<size_t>(feature_len(c, RML_MASK_ALL)) * 4;
    const size_t per = net_tower_bytes_per_scan(c) + F4;
    if (ws_bytes > flat + 4096 + per) {
      int64_t chunk = static_cast<int64_t>((ws_bytes - flat - 4096) / per);
      if (chunk > group) chunk = group;
      if (chunk >= 1) {
        pl->chunk = chunk; pl->group = group;
        pl->flat_off = align256(net_tower_bytes_per_scan(c) * chunk + 1024);
        pl->feats_off = pl->flat_off + flat;
        return RML_OK;
      }
    }
    if (group <= 128) break;
    group /= 2;
  }
  return fail(c, RML_E_INVALID, "network workspace too small: %zu bytes", ws_bytes);
}

int rml_net_uses_igemm(const rml_ctx* c) { return c && c->net.ready && c->net.use_igemm; }

size_t rml_net_workspace_bytes(const rml_ctx* c, int64_t chunk) {
  if (!c || !c->net.ready || chunk <= 0) return 0;
  return net_plan_bytes(c, chunk, kDenseGroup);
}

static int net_resize(rml_ctx* c, const float* feats, int64_t n_scans, float* images, cudaStream_t st) {
  Net& n = c->net;
  ResizeParams rp;
  rp.feats = feats; rp.images = images; rp.B = n_scans; rp.F = feature_len(c, RML_MASK_ALL); rp.R = n.R;
  const int ph[3] = {c->sx, c->sy, c->sx}, pw[3] = {c->sz, c->sz, c->sy};
  const int poff[3] = {0, c->sx * c->sz, c->sx * c->sz + c->sy * c->sz};
  int smem_max = 0;
  for (int b = 0; b < 3; ++b) {
    rp.ph[b] = ph[b]; rp.pw[b] = pw[b]; rp.poff[b] = poff[b];
    rp.kh[b] = n.kh[b]; rp.kv[b] = n.kv[b]; rp.bh[b] = n.bh[b]; rp.bv[b] = n.bv[b];
    rp.ksh[b] = n.ksh[b]; rp.ksv[b] = n.ksv[b];
    const int sm = (ph[b] * pw[b] + ph[b] * n.R) * 4;
    if (sm > smem_max) smem_max = sm;
  }
  const int gx = static_cast<int>(n_scans < 4ll * c->num_sms ? n_scans : 4ll * c->num_sms);
  k3_resize_pil<<<dim3(gx, 3), 256, smem_max, st>>>(rp);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

// Towers of one chunk -> flat (bf16 [n_scans][K]).  feats != null: K3 resize into the workspace
// first; else images_in [n][3][R][R] is used as is.
static int net_towers_chunk(rml_ctx* c, const float* feats, const float* images_in, int64_t n_scans,
                            char* ws, uint16_t* flat, cudaStream_t st) {
  Net& n = c->net;
  size_t img1, act;
  net_tower_sizes(n, &img1, &act);
  size_t img_b = align256(img1 * static_cast<size_t>(n_scans));
  size_t act_b = align256(act * static_cast<size_t>(n_scans));
  float* images = reinterpret_cast<float*>(ws);
  float* ping = reinterpret_cast<float*>(ws + img_b);
  float* pong = reinterpret_cast<float*>(ws + img_b + act_b);
  if (n.tower_mode) {
    TowerParams tp;
    tp.images = feats ? nullptr : images_in;
    tp.rz.feats = feats; tp.rz.images = nullptr; tp.rz.B = n_scans; tp.rz.F = feature_len(c, RML_MASK_ALL); tp.rz.R = n.R;
    const int ph[3] = {c->sx, c->sy, c->sx}, pw[3] = {c->sz, c->sz, c->sy};
    const int poff[3] = {0, c->sx * c->sz, c->sx * c->sz + c->sy * c->sz};
    for (int b = 0; b < 3; ++b) {
      tp.rz.ph[b] = ph[b]; tp.rz.pw[b] = pw[b]; tp.rz.poff[b] = poff[b];
      tp.rz.kh[b] = n.kh[b]; tp.rz.kv[b] = n.kv[b]; tp.rz.bh[b] = n.bh[b]; tp.rz.bv[b] = n.bv[b];
      tp.rz.ksh[b] = n.ksh[b]; tp.rz.ksv[b] = n.ksv[b];
      tp.ctas[b] = n.t6_ctas[b];
    }
    tp.B = n_scans;
    tp.w1 = reinterpret_cast<const __half*>(n.t6_w1);
    tp.alpha = n.alpha;
    tp.dbg = n.t6_dbg;
    const int grid = n.t6_ctas[0] + n.t6_ctas[1] + n.t6_ctas[2];
    if (n.tower_mode == 1) {
      tp.w2 = reinterpret_cast<const __nv_bfloat16*>(n.convs[1].wt_bf16); tp.b2 = n.convs[1].bias3;
      tp.out = reinterpret_cast<__nv_bfloat16*>(flat);
      k6_tower<64, true, 1><<<grid, kT6Threads, T6Smem<64, true>::total, st>>>(tp, CUtensorMap{});
      RML_CUDA(c, cudaGetLastError());
      ++c->launches;
      return RML_OK;
    }
    tp.w2 = nullptr; tp.b2 = nullptr;
    tp.out = reinterpret_cast<__nv_bfloat16*>(ping);
    // layer-1 rows leave through tiled TMA stores: [pixels][128 ch] bf16, box {64 ch, 32 px}, 128B swizzle
    CUtensorMap map_out;
    {
      const int h1 = n.R / 2;
      const int64_t pixels = n_scans * 3 * h1 * h1;
      if (pixels > 0x7fffffffll) return fail(c, RML_E_INVALID, "tower chunk too large: %lld layer-1 pixels", static_cast<long long>(pixels));
      cuuint64_t gdim[2] = {128u, static_cast<cuuint64_t>(pixels)};
      cuuint64_t gstr[1] = {256u};
      cuuint32_t box[2] = {64u, 32u};
      cuuint32_t estr[2] = {1u, 1u};
      CUresult r = c->encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ping, gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(c, RML_E_CUDA, "cuTensorMapEncodeTiled(tower output) failed: %d", (int)r);
    }
    k6_tower<128, false, 2><<<grid, kT6Threads, T6Smem<128, false>::total, st>>>(tp, map_out);
    RML_CUDA(c, cudaGetLastError());
    ++c->launches;
  }
  const bool fuse1 = feats && !n.tower_mode && n.use_igemm && n.fuse_resize && (n.convs[0].cout == 64 || n.convs[0].cout == 128);
  if (feats && !fuse1 && !n.tower_mode) {
    int rc = net_resize(c, feats, n_scans, images, st);
    if (rc) return rc;
  }
  // K4 conv towers
  const void* cur = feats ? images : images_in;
  int hw = n.R;
  size_t l0 = 0;
  if (n.tower_mode == 2) {
    cur = ping;
    hw = (hw + 1) / 2;
    l0 = 1;
  }
  if (fuse1) {
    // K3 + first tower layer in one kernel: the resized image never leaves shared memory
    const NetConv& cv = n.convs[0];
    ResizeConv1Params fp;
    fp.rz.feats = feats; fp.rz.images = nullptr; fp.rz.B = n_scans; fp.rz.F = feature_len(c, RML_MASK_ALL); fp.rz.R = n.R;
    const int ph[3] = {c->sx, c->sy, c->sx}, pw[3] = {c->sz, c->sz, c->sy};
    const int poff[3] = {0, c->sx * c->sz, c->sx * c->sz + c->sy * c->sz};
    int smem_max = 0;
    for (int b = 0; b < 3; ++b) {
      fp.rz.ph[b] = ph[b]; fp.rz.pw[b] = pw[b]; fp.rz.poff[b] = poff[b];
      fp.rz.kh[b] = n.kh[b]; fp.rz.kv[b] = n.kv[b]; fp.rz.bh[b] = n.bh[b]; fp.rz.bv[b] = n.bv[b];
      fp.rz.ksh[b] = n.ksh[b]; fp.rz.ksv[b] = n.ksv[b];
      fp.w[b] = cv.w[b]; fp.bias[b] = cv.bias[b];
      const int sm = k34_smem_floats(ph[b], pw[b], n.R) * 4;
      if (sm > smem_max) smem_max = sm;
    }
    const int ho = (hw + 1) / 2;
    const int pad_total = (ho - 1) * 2 + 3 - hw;
    fp.Cout = cv.cout; fp.Ho = ho; fp.Wo = ho; fp.pad_t = fp.pad_l = pad_total > 0 ? pad_total / 2 : 0;
    fp.act = cv.act; fp.alpha = n.alpha;
    fp.out = reinterpret_cast<__nv_bfloat16*>(ping);
    const int gx = static_cast<int>(n_scans < 8ll * c->num_sms ? n_scans : 8ll * c->num_sms);
    if (cv.cout == 64) k34_resize_conv1<2><<<dim3(gx, 3), 256, smem_max, st>>>(fp);
    else k34_resize_conv1<4><<<dim3(gx, 3), 256, smem_max, st>>>(fp);
    RML_CUDA(c, cudaGetLastError());
    ++c->launches;
    cur = ping;
    hw = ho;
    l0 = 1;
  }
  for (size_t l = l0; l < n.convs.size(); ++l) {
    const NetConv& cv = n.convs[l];
    const bool last = l + 1 == n.convs.size();
    const int ho = (hw + 1) / 2;
    const int pad_total = (ho - 1) * 2 + 3 - hw;
    const int pad = pad_total > 0 ? pad_total / 2 : 0;       // TF 'same': extra padding goes after
    void* dst = last ? static_cast<void*>(flat) : static_cast<void*>((l & 1) ? pong : ping);
    if (l >= 1 && n.use_igemm) {
      // tcgen05 implicit GEMM on the NHWC bf16 activation written by the previous layer
      CUtensorMap map_x;
      cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cv.cin), static_cast<cuuint64_t>(hw),
                            static_cast<cuuint64_t>(hw), static_cast<cuuint64_t>(n_scans * 3)};
      cuuint64_t gstr[3] = {static_cast<cuuint64_t>(cv.cin) * 2, static_cast<cuuint64_t>(hw) * cv.cin * 2,
                            static_cast<cuuint64_t>(hw) * hw * cv.cin * 2};
      int th = 128 / ho;
      if (th > ho) th = ho;
      cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(2 * ho), static_cast<cuuint32_t>(2 * th), 1u};
      cuuint32_t estr[4] = {1u, 2u, 2u, 1u};
      CUresult r = c->encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(cur), gdim, gstr,
                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(c, RML_E_CUDA, "cuTensorMapEncodeTiled(conv activation) failed: %d", (int)r);
      ConvGemmParams gp;
      gp.n_img = n_scans * 3; gp.Ho = ho; gp.Wo = ho; gp.Cin = cv.cin; gp.Cout = cv.cout; gp.TH = th;
      gp.tiles_per_img = (ho + th - 1) / th; gp.pad_t = pad; gp.pad_l = pad;
      gp.act = cv.act; gp.alpha = n.alpha; gp.bias = cv.bias3;
      gp.out = reinterpret_cast<__nv_bfloat16*>(dst);
      // taps kh = 0 / kh = 2 from one box of TH + 1 even rows (see ConvGemmParams::share_kh)
      gp.share_kh = (ho % 8 == 0 && ho <= 32 && th * ho <= 128 && ho % th == 0 && n.k4_share) ? 1 : 0;
      CUtensorMap map_xe = map_x;
      if (gp.share_kh) {
        cuuint32_t box_e[4] = {64u, static_cast<cuuint32_t>(2 * ho), static_cast<cuuint32_t>(2 * (th + 1)), 1u};
        r = c->encode(&map_xe, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(cur), gdim, gstr,
                      box_e, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, RML_E_CUDA, "cuTensorMapEncodeTiled(conv activation, even rows) failed: %d", (int)r);
      }
      gp.stages = cg_pick_stages(cv.cout, gp.share_kh);
      const int smem = cg_smem_bytes(cv.cout, gp.stages, gp.share_kh);
      const int64_t tiles = gp.n_img * gp.tiles_per_img;
      const int grid = static_cast<int>(tiles < c->num_sms ? tiles : c->num_sms);
      k4_conv_igemm<<<grid, kCgThreads, smem, st>>>(map_x, cv.map_w, map_xe, gp);
    } else if (cv.cin == 1 && n.use_igemm && (cv.cout == 64 || cv.cout == 128)) {
      Conv1Params c1;
      c1.in = static_cast<const float*>(cur);
      c1.out = reinterpret_cast<__nv_bfloat16*>(dst);
      for (int b = 0; b < 3; ++b) { c1.w[b] = cv.w[b]; c1.bias[b] = cv.bias[b]; }
      c1.n_img = n_scans * 3; c1.H = hw; c1.W = hw; c1.Cout = cv.cout; c1.Ho = ho; c1.Wo = ho;
      c1.pad_t = c1.pad_l = pad; c1.act = cv.act; c1.alpha = n.alpha;
      const int64_t tasks = c1.n_img * ho;     // one warp task per output row
      int64_t blocks = (tasks + 7) / 8;
      if (blocks > 16ll * c->num_sms) blocks = 16ll * c->num_sms;
      if (cv.cout == 64) k4_conv1_cin1<2><<<static_cast<unsigned>(blocks), 256, 0, st>>>(c1);
      else k4_conv1_cin1<4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(c1);
    } else {
      ConvParams cp;
      cp.in = static_cast<const float*>(cur);
      cp.out = dst;
      for (int b = 0; b < 3; ++b) { cp.w[b] = cv.w[b]; cp.bias[b] = cv.bias[b]; }
      cp.n_img = n_scans * 3; cp.H = hw; cp.W = hw; cp.Cin = cv.cin; cp.Cout = cv.cout;
      cp.Ho = ho; cp.Wo = ho; cp.pad_t = cp.pad_l = pad;
      cp.act = cv.act; cp.alpha = n.alpha;
      cp.out_bf16 = (last || n.use_igemm) ? 1 : 0;           // bf16 feeds the tensor-core layers
      const int smem = 9 * cv.cin * kConvCoutTile * 4;
      const int64_t pixels = n_scans * ho * ho;
      int64_t gx = (pixels + 255) / 256;
      if (gx > 8ll * c->num_sms) gx = 8ll * c->num_sms;
      k4_conv3x3s2<<<dim3(static_cast<unsigned>(gx), cv.cout / kConvCoutTile, 3), 256, smem, st>>>(cp);
    }
    RML_CUDA(c, cudaGetLastError());
    ++c->launches;
    cur = dst;
    hw = ho;
  }
  return RML_OK;
}

// K5 dense stack on tcgen05 over a whole dense group
static int net_dense(rml_ctx* c, const uint16_t* flat, int64_t n_scans, float* proba, float* logits,
                     int32_t* label, cudaStream_t st) {
  Net& n = c->net;
  CUtensorMap map_act;
  int rc = encode_bf16_map(c, &map_act, flat, n_scans, n.K, kK5BlockM);
  if (rc) return rc;
  K5Params kp;
  kp.B = n_scans; kp.k_blocks = n.K / kK5BlockK; kp.C = n.C; kp.head = n.head;
  kp.kpg = (kp.k_blocks % 4 == 0) ? 4 : (kp.k_blocks % 2 == 0) ? 2 : 1;
  if (c->k5_kpg && kp.k_blocks % c->k5_kpg == 0) kp.kpg = c->k5_kpg;
  kp.act1 = n.act1; kp.act2 = n.act2; kp.alpha = n.alpha;
  kp.b1 = n.b1; kp.w2 = n.w2; kp.b2 = n.b2; kp.w3 = n.w3; kp.b3 = n.b3;
  kp.proba = proba; kp.logits = logits; kp.label = label;
  const int smem5 = k5_smem_bytes();
  const int64_t tiles = (n_scans + kK5BlockM - 1) / kK5BlockM;
  const int grid = static_cast<int>(tiles < c->num_sms ? tiles : c->num_sms);
  k5_dense_stack<<<grid, kK5Threads, smem5, st>>>(map_act, n.map_w1, kp);
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

// shared driver: conv towers chunk by chunk, dense stack once per dense group
static int net_run(rml_ctx* c, const void* cubes, int mode, const int32_t* ijk, const float* feats,
                   const float* images, int64_t B, void* workspace, size_t workspace_bytes, float* proba,
                   float* logits, int32_t* label, uint16_t* tower_bf16, cudaStream_t st, int cube_u8 = 0) {
  NetPlan pl;
  int rc = net_make_plan(c, B, workspace_bytes, &pl);
  if (rc) return rc;
  char* ws = static_cast<char*>(workspace);
  uint16_t* flat = reinterpret_cast<uint16_t*>(ws + pl.flat_off);
  float* feats_ws = reinterpret_cast<float*>(ws + pl.feats_off);
  const int F = feature_len(c, RML_MASK_ALL);
  const int C = c->net.C;
  const int64_t K = c->net.K;
  const size_t cube_elems = static_cast<size_t>(c->sx) * c->sy * c->sz;
  const size_t img_elems = static_cast<size_t>(3) * c->net.R * c->net.R;
  for (int64_t g0 = 0; g0 < B; g0 += pl.group) {
    const int64_t gn = (B - g0) < pl.group ? (B - g0) : pl.group;
    for (int64_t lo = 0; lo < gn; lo += pl.chunk) {
      const int64_t n = (gn - lo) < pl.chunk ? (gn - lo) : pl.chunk;
      const int64_t s0 = g0 + lo;
      const float* f = nullptr;
      const float* im = nullptr;
      if (cubes) {
        // K1 with the network's scaling (p - 127.5) / 127.5 (dnn.py:202-205)
        const Affine aff{127.5f, 127.5f, 1};
        const char* cube0 = static_cast<const char*>(cubes) + s0 * cube_elems * (cube_u8 ? 1 : 4);
        rc = project_impl(c, cube0, n, mode, ijk ? ijk + s0 * 3 : nullptr, RML_MASK_ALL,
                          RML_F32, feats_ws, nullptr, st, 0, nullptr, &aff, cube_u8);
        if (rc) return rc;
        f = feats_ws;
      } else if (feats) {
        f = feats + s0 * F;
      } else {
        im = images + s0 * img_elems;
      }
      rc = net_towers_chunk(c, f, im, n, ws, flat + lo * K, st);
      if (rc) return rc;
    }
    if (tower_bf16)
      RML_CUDA(c, cudaMemcpyAsync(tower_bf16 + g0 * K, flat, static_cast<size_t>(gn) * K * 2,
                                  cudaMemcpyDeviceToDevice, st));
    rc = net_dense(c, flat, gn, proba + g0 * C, logits ? logits + g0 * C : nullptr, label + g0, st);
    if (rc) return rc;
  }
  return RML_OK;
}

int rml_net_forward(rml_ctx* c, const float* feats, int64_t B, void* workspace, size_t workspace_bytes,
                    float* proba, float* logits, int32_t* label, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!c->net.ready) return fail(c, RML_E_NOMODEL, "rml_net_forward: no network loaded");
  if (!feats || !workspace || !proba || !label || B < 0) return fail(c, RML_E_INVALID, "rml_net_forward: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  return net_run(c, nullptr, 0, nullptr, feats, nullptr, B, workspace, workspace_bytes, proba, logits, label,
                 nullptr, static_cast<cudaStream_t>(stream));
}

// dnn.py:240-254 alone: scaled projections -> [B][3][R][R] float32 (the Keras model's inputs)
int rml_net_resize(rml_ctx* c, const float* feats, int64_t B, float* images, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!c->net.kh[0] || !c->net.kh[1] || !c->net.kh[2]) return fail(c, RML_E_NOMODEL, "rml_net_resize: resize tables not loaded");
  if (!feats || !images || B < 0) return fail(c, RML_E_INVALID, "rml_net_resize: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  return net_resize(c, feats, B, images, static_cast<cudaStream_t>(stream));
}

// Keras model.predict([XZ, YZ, XY]) on already preprocessed inputs: images [B][3][R][R]
int rml_net_forward_images(rml_ctx* c, const float* images, int64_t B, void* workspace,
                           size_t workspace_bytes, float* proba, float* logits, int32_t* label,
                           uint16_t* tower_bf16, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!c->net.ready) return fail(c, RML_E_NOMODEL, "rml_net_forward_images: no network loaded");
  if (!images || !workspace || !proba || !label || B < 0) return fail(c, RML_E_INVALID, "rml_net_forward_images: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  return net_run(c, nullptr, 0, nullptr, nullptr, images, B, workspace, workspace_bytes, proba, logits, label,
                 tower_bf16, static_cast<cudaStream_t>(stream));
}

// cubes -> K1 (projection + (p-127.5)/127.5, dnn.py:202-205) -> K3 -> K4 -> K5
int rml_net_predict(rml_ctx* c, const float* cubes, int64_t B, int mode, const int32_t* ijk,
                    void* workspace, size_t workspace_bytes, float* proba, int32_t* label,
                    rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!c->net.ready) return fail(c, RML_E_NOMODEL, "rml_net_predict: no network loaded");
  if (!cubes || !workspace || !proba || !label || B < 0) return fail(c, RML_E_INVALID, "rml_net_predict: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  return net_run(c, cubes, mode, ijk, nullptr, nullptr, B, workspace, workspace_bytes, proba, nullptr, label,
                 nullptr, static_cast<cudaStream_t>(stream));
}

int rml_net_predict_u8(rml_ctx* c, const uint8_t* cubes, int64_t B, int mode, const int32_t* ijk,
                       void* workspace, size_t workspace_bytes, float* proba, int32_t* label,
                       rml_stream stream) {
  if (!c) return RML_E_INVALID;
  if (!c->net.ready) return fail(c, RML_E_NOMODEL, "rml_net_predict_u8: no network loaded");
  if (!cubes || !workspace || !proba || !label || B < 0) return fail(c, RML_E_INVALID, "rml_net_predict_u8: null buffer or B<0");
  if (B == 0) return RML_OK;
  DeviceGuard g(c->device);
  return net_run(c, cubes, mode, ijk, nullptr, nullptr, B, workspace, workspace_bytes, proba, nullptr, label,
                 nullptr, static_cast<cudaStream_t>(stream), 1);
}

}  // extern "C"
