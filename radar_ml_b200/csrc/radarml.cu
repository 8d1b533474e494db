// libradarml.so — C ABI (include/radarml.h) over the sm_100a kernels.
// Host side: context + model residency, tensor-map encoding, launches, host-buffer pipeline.
#include "../../include/radarml.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include "k1_project.cuh"
#include "k2_score.cuh"

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
  double* rho = nullptr;      // svc: [NP]; linear: intercept [R]
  double* platt_a = nullptr;
  double* platt_b = nullptr;
  int kpad = 0, n_tile = 0, n_chunks = 0;
  CUtensorMap map_sv;
};

constexpr int kHostBufs = 3;
struct HostPipe {
  int64_t chunk = 0;
  float* cubes[kHostBufs] = {nullptr};
  int32_t* ijk[kHostBufs] = {nullptr};
  void* work[kHostBufs] = {nullptr};
  float* proba[kHostBufs] = {nullptr};
  int32_t* label[kHostBufs] = {nullptr};
  uint8_t* known[kHostBufs] = {nullptr};
  cudaStream_t stream[kHostBufs] = {nullptr};
  size_t work_bytes = 0;
};

}  // namespace

struct rml_ctx {
  int device = 0;
  int num_sms = 148;
  int sx = kSX, sy = kSY, sz = kSZ;
  double r_min = 10, r_max = 360, th_min = -42, th_max = 42, ph_min = -30, ph_max = 30;
  float aff_offset = 0.f, aff_scale = 255.f;
  int aff_enabled = 1;
  Model model;
  unsigned int* status = nullptr;  // [0] non-integral values, [1] slice index errors
  int64_t launches = 0;
  std::string err;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  HostPipe pipe;
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
  cudaFree(m.rho);
  cudaFree(m.platt_a);
  cudaFree(m.platt_b);
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
int launch_rbf_i8(rml_ctx* c, const CUtensorMap& map_feats, const K2Params& p, cudaStream_t st) {
  const int smem = k2_smem_bytes(p.n_tile, p.stages);
  RML_CUDA(c, cudaFuncSetAttribute(k2_rbf_i8<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t tiles = (p.B + kK2BlockM - 1) / kK2BlockM;
  const int grid = static_cast<int>(tiles < c->num_sms ? tiles : c->num_sms);
  k2_rbf_i8<C><<<grid, kK2Threads, smem, st>>>(map_feats, c->model.map_sv, p);
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

int project_impl(rml_ctx* c, const float* cubes, int64_t B, int mode, const int32_t* ijk,
                 uint32_t mask, int dtype, void* feats, int32_t* norms, cudaStream_t st) {
  if (B < 0 || !cubes || !feats) return fail(c, RML_E_INVALID, "rml_project: null buffer or B<0");
  if ((mask & RML_MASK_ALL) == 0 || (mask & ~RML_MASK_ALL))
    return fail(c, RML_E_INVALID, "rml_project: mask %u selects no projection", mask);
  if (mode != RML_MODE_MAX && mode != RML_MODE_SLICE) return fail(c, RML_E_INVALID, "bad mode %d", mode);
  if (mode == RML_MODE_SLICE && !ijk) return fail(c, RML_E_INVALID, "SLICE mode needs ijk");
  if (dtype != RML_F32 && dtype != RML_U8) return fail(c, RML_E_INVALID, "bad dtype %d", dtype);
  if ((reinterpret_cast<uintptr_t>(cubes) & 15) || (reinterpret_cast<uintptr_t>(feats) & 15))
    return fail(c, RML_E_INVALID, "cubes/feats must be 16-byte aligned");
  if (B == 0) return RML_OK;
  const int F = feature_len(c, mask);
  const int stride = feature_stride(c, mask, dtype);
  const bool fast = mode == RML_MODE_MAX && c->sx == kSX && c->sy == kSY && c->sz == kSZ;
  if (fast) {
    K1Params p;
    p.cubes = cubes; p.feats = feats; p.norms = norms; p.status = c->status; p.B = B;
    p.stride = stride; p.F = F; p.mask = mask;
    p.offset = c->aff_offset; p.scale = c->aff_scale; p.affine = c->aff_enabled;
    const int grid = static_cast<int>(B < c->num_sms ? B : c->num_sms);
    if (dtype == RML_U8) {
      const int smem = k1_smem_bytes<uint8_t>();
      RML_CUDA(c, cudaFuncSetAttribute(k1_project_max<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      k1_project_max<uint8_t><<<grid, kK1Threads, smem, st>>>(p);
    } else {
      const int smem = k1_smem_bytes<float>();
      RML_CUDA(c, cudaFuncSetAttribute(k1_project_max<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      k1_project_max<float><<<grid, kK1Threads, smem, st>>>(p);
    }
  } else {
    K1GenParams p;
    p.cubes = cubes; p.ijk = ijk; p.feats = feats; p.norms = norms; p.status = c->status; p.B = B;
    p.sx = c->sx; p.sy = c->sy; p.sz = c->sz; p.stride = stride; p.F = F; p.mask = mask;
    p.offset = c->aff_offset; p.scale = c->aff_scale; p.affine = c->aff_enabled; p.mode = mode;
    const int64_t want = B < 8ll * c->num_sms ? B : 8ll * c->num_sms;
    const int grid = static_cast<int>(want);
    if (dtype == RML_U8) k1_project_generic<uint8_t><<<grid, 256, 0, st>>>(p);
    else k1_project_generic<float><<<grid, 256, 0, st>>>(p);
  }
  RML_CUDA(c, cudaGetLastError());
  ++c->launches;
  return RML_OK;
}

int score_impl(rml_ctx* c, const void* feats, int dtype, const int32_t* norms, int64_t B,
               double min_proba, float* proba, float* decision, int32_t* label, uint8_t* known,
               cudaStream_t st) {
  Model& m = c->model;
  if (m.kind == 0) return fail(c, RML_E_NOMODEL, "rml_score: no model loaded");
  if (!feats || !proba || !label || B < 0) return fail(c, RML_E_INVALID, "rml_score: null buffer or B<0");
  if (B == 0) return RML_OK;
  if (m.kind == 2) {
    K2LinParams p;
    p.B = B; p.F = m.F; p.dtype = dtype;
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
    p.stages = k2_pick_stages(m.n_tile);
    p.unorm = norms; p.svnorm = m.svnorm; p.coef = m.coef; p.rho = m.rho;
    p.platt_a = m.platt_a; p.platt_b = m.platt_b;
    p.neg_gamma_s2 = -m.gamma / (m.feature_scale * m.feature_scale);
    p.min_proba = min_proba; p.proba = proba; p.decision = decision; p.label = label; p.known = known;
    for (int i = 0; i < kMaxClasses; ++i) p.class_end[i] = m.class_end[i];
    DISPATCH_C(m.C, launch_rbf_i8<CC>(c, map_feats, p, st));
  }
  K2GenParams p;
  p.B = B; p.F = m.F; p.n_sv = m.n_sv; p.feats = static_cast<const float*>(feats); p.sv = m.sv_f64;
  p.coef = m.coef; p.rho = m.rho; p.platt_a = m.platt_a; p.platt_b = m.platt_b;
  p.neg_gamma = -m.gamma; p.min_proba = min_proba;
  p.proba = proba; p.decision = decision; p.label = label; p.known = known;
  for (int i = 0; i < kMaxClasses; ++i) p.class_end[i] = m.class_end[i];
  DISPATCH_C(m.C, launch_rbf_general<CC>(c, p, st));
}

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

bool use_u8_path(const rml_ctx* c) {
  return c->model.kind == 2 || (c->model.kind == 1 && c->model.integral);
}

int predict_impl(rml_ctx* c, const float* cubes, int64_t B, int mode, const int32_t* ijk,
                 uint32_t mask, double min_proba, void* work, float* proba, int32_t* label,
                 uint8_t* known, cudaStream_t st) {
  if (c->model.kind == 0) return fail(c, RML_E_NOMODEL, "rml_predict: no model loaded");
  if (feature_len(c, mask) != c->model.F)
    return fail(c, RML_E_INVALID, "rml_predict: mask gives F=%d but the model has F=%d",
                feature_len(c, mask), c->model.F);
  if (!work) return fail(c, RML_E_INVALID, "rml_predict: workspace is null");
  const int dtype = use_u8_path(c) ? RML_U8 : RML_F32;
  const size_t stride = feature_stride(c, mask, dtype);
  const size_t feat_bytes = align256(static_cast<size_t>(B) * stride * (dtype == RML_U8 ? 1 : 4));
  int32_t* norms = reinterpret_cast<int32_t*>(static_cast<char*>(work) + feat_bytes);
  const int saved = c->aff_enabled;
  const float so = c->aff_offset, ss = c->aff_scale;
  // the scorer expects features scaled like common.process_samples(scale=True)
  c->aff_enabled = 1; c->aff_offset = 0.f; c->aff_scale = static_cast<float>(c->model.feature_scale);
  int rc = project_impl(c, cubes, B, mode, ijk, mask, dtype, work, norms, st);
  c->aff_enabled = saved; c->aff_offset = so; c->aff_scale = ss;
  if (rc) return rc;
  return score_impl(c, work, dtype, norms, B, min_proba, proba, nullptr, label, known, st);
}

void free_pipe(HostPipe& hp) {
  for (int i = 0; i < kHostBufs; ++i) {
    cudaFree(hp.cubes[i]); cudaFree(hp.ijk[i]); cudaFree(hp.work[i]);
    cudaFree(hp.proba[i]); cudaFree(hp.label[i]); cudaFree(hp.known[i]);
    if (hp.stream[i]) cudaStreamDestroy(hp.stream[i]);
  }
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
  *out = c;
  return RML_OK;
}

int rml_destroy(rml_ctx* c) {
  if (!c) return RML_OK;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  free_model(c->model);
  free_pipe(c->pipe);
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
  c->aff_offset = offset; c->aff_scale = scale; c->aff_enabled = enabled;
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
  std::vector<int32_t> norm(n_sv, 0);
  bool integral = true;
  for (int n = 0; n < n_sv && integral; ++n) {
    int64_t s2 = 0;
    for (int f = 0; f < F; ++f) {
      const double v = sv[static_cast<size_t>(n) * F + f] * feature_scale;
      const double r = std::nearbyint(v);
      if (std::fabs(v - r) > 1e-4 || r < 0 || r > 255) { integral = false; break; }
      u8[static_cast<size_t>(n) * m.kpad + f] = static_cast<uint8_t>(r);
      s2 += static_cast<int64_t>(r) * static_cast<int64_t>(r);
    }
    norm[n] = static_cast<int32_t>(s2);
  }
  // 2*max(u.s) and the norm sum must stay inside s32: F * 255^2 * 2 < 2^31  <=>  F <= 16512
  if (static_cast<double>(F) * 65025.0 * 2.0 >= 2147483647.0) integral = false;
  m.integral = integral;
  int rc;
  if (integral) {
    if ((rc = upload(c, &m.sv_u8, u8.data(), u8.size()))) return rc;
    if ((rc = upload(c, &m.svnorm, norm.data(), norm.size()))) return rc;
    if ((rc = encode_u8_map(c, &m.map_sv, m.sv_u8, n_sv, F, m.kpad, m.n_tile))) return rc;
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
  p.feats = feats; p.B = B; p.F = feature_len(c, mask); p.scale = scale; p.scale_value = 255.f;
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
  const int dtype = use_u8_path(c) ? RML_U8 : RML_F32;
  const size_t stride = feature_stride(c, RML_MASK_ALL, dtype);
  return align256(static_cast<size_t>(B) * stride * (dtype == RML_U8 ? 1 : 4)) +
         align256(static_cast<size_t>(B) * 4) + 256;
}

int rml_predict(rml_ctx* c, const float* cubes, int64_t B, int mode, const int32_t* ijk,
                uint32_t mask, double min_proba, void* work, float* proba, int32_t* label,
                uint8_t* known, rml_stream stream) {
  if (!c) return RML_E_INVALID;
  DeviceGuard g(c->device);
  return predict_impl(c, cubes, B, mode, ijk, mask, min_proba, work, proba, label, known,
                      static_cast<cudaStream_t>(stream));
}

int rml_predict_host(rml_ctx* c, const float* cubes_host, int64_t B, int mode,
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
  const int64_t chunk = 512;  // 512 cubes = 246 MB per H2D transfer
  const size_t cube_elems = static_cast<size_t>(c->sx) * c->sy * c->sz;
  const int C = c->model.C;
  if (hp.chunk != chunk) {
    free_pipe(hp);
    hp.work_bytes = rml_predict_workspace_bytes(c, chunk);
    for (int i = 0; i < kHostBufs; ++i) {
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.cubes[i]), chunk * cube_elems * 4));
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.ijk[i]), chunk * 3 * 4));
      RML_CUDA(c, cudaMalloc(&hp.work[i], hp.work_bytes));
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.proba[i]), chunk * kMaxClasses * 4));
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.label[i]), chunk * 4));
      RML_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&hp.known[i]), chunk));
      RML_CUDA(c, cudaStreamCreateWithFlags(&hp.stream[i], cudaStreamNonBlocking));
    }
    hp.chunk = chunk;
  }
  int64_t done = 0;
  int slot = 0;
  while (done < B) {
    const int64_t n = (B - done) < chunk ? (B - done) : chunk;
    cudaStream_t st = hp.stream[slot];
    RML_CUDA(c, cudaMemcpyAsync(hp.cubes[slot], cubes_host + done * cube_elems, n * cube_elems * 4,
                                cudaMemcpyHostToDevice, st));
    if (mode == RML_MODE_SLICE)
      RML_CUDA(c, cudaMemcpyAsync(hp.ijk[slot], ijk_host + done * 3, n * 12, cudaMemcpyHostToDevice, st));
    int rc = predict_impl(c, hp.cubes[slot], n, mode, hp.ijk[slot], mask, min_proba, hp.work[slot],
                          hp.proba[slot], hp.label[slot], hp.known[slot], st);
    if (rc) return rc;
    RML_CUDA(c, cudaMemcpyAsync(proba_host + done * C, hp.proba[slot], n * C * 4, cudaMemcpyDeviceToHost, st));
    RML_CUDA(c, cudaMemcpyAsync(label_host + done, hp.label[slot], n * 4, cudaMemcpyDeviceToHost, st));
    if (known_host)
      RML_CUDA(c, cudaMemcpyAsync(known_host + done, hp.known[slot], n, cudaMemcpyDeviceToHost, st));
    done += n;
    slot = (slot + 1) % kHostBufs;
  }
  for (int i = 0; i < kHostBufs; ++i) RML_CUDA(c, cudaStreamSynchronize(hp.stream[i]));
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
  if (h[1]) return fail(c, RML_E_INVALID, "SLICE mode: %u scans had a target index outside the cube (numpy IndexError)", h[1]);
  if (h[0]) return fail(c, RML_E_NONINTEGRAL, "u8 path: %u warps saw values that are not integers in [0,255]; use RML_F32", h[0]);
  return RML_OK;
}

int64_t rml_launch_count(const rml_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
