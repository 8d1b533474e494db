// K2 rbf_platt_argmax — replaces predict.py:56-70 classifier(): model.predict_proba on the
// CalibratedClassifierCV(prefit SVC-RBF) built at train.py:478-479, 723-724, i.e.
//   libsvm k_function RBF          SK/svm/src/libsvm/svm.cpp:461-514
//   svm_predict_values OvO sums    svm.cpp:2864-2893
//   _ovr_decision_function         SK/utils/multiclass.py:557-599 (called with dec<0, -dec)
//   Platt expit(-(a f + b))        SK/calibration.py:1065
//   row normalise / uniform / clip SK/calibration.py:824-847
//   argmax + >= min_proba          predict.py:61-68
//
// Integer-exact tensor-core form: real sensor values and un-augmented support vectors are
// integers u,s in [0,255] (features = u/255), so
//   ||x - sv||^2 = (||u||^2 + ||s||^2 - 2 u.s) / 255^2
// with u.s accumulated EXACTLY in s32 by tcgen05.mma kind::i8 (u8 x u8 -> s32, max 6.5e8).
// Warp-specialised: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue
// (TMEM -> registers, fp64 exp / dual-coef sums / OvR / Platt / argmax).  Accumulators are
// double-buffered in TMEM (2 x 256 columns) so the epilogue of one support-vector chunk
// overlaps the MMAs of the next.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace rml {

constexpr int kK2EpiWarps = 8;                 // two warps per TMEM lane quarter
constexpr int kK2Threads = (2 + kK2EpiWarps) * 32;   // 320
constexpr int kK2BlockM = 128;
constexpr int kK2BlockKBytes = 128;          // one 128-byte swizzle span of u8 features
constexpr int kK2MaxTileN = 256;             // columns per accumulator buffer
constexpr int kK2MaxStages = 8;
constexpr int kMaxClasses = 8;

struct K2Params {
  int64_t B;
  int n_sv;
  int n_tile;     // support vectors per chunk (multiple of 16, <= 256)
  int n_chunks;
  int k_blocks;   // padded feature bytes / 128
  int stages;     // smem ring depth (host picks the deepest that fits 227 KB)
  const int32_t* unorm;    // [B]    sum u^2
  const int32_t* svnorm;   // [n_pad] sum s^2 (n_pad = n_tile * n_chunks, zero padded)
  const double* pairw;     // [C(C-1)/2][n_pad] weight of SV n in OvO pair p (0 outside the pair)
  int n_pad;
  const unsigned int* tile_ready;  // nullable: [tiles] scans of the tile finished by the producer kernel
  const double* rho;       // [C(C-1)/2]
  const double* platt_a;   // [C] (or [1] when C == 2)
  const double* platt_b;
  double neg_gamma_s2;     // -gamma / feature_scale^2
  double min_proba;
  float* proba;            // [B][C]
  float* decision;         // [B][C] or [B] (C == 2), nullable
  int32_t* label;          // [B]
  uint8_t* known;          // [B], nullable
};

__host__ __device__ constexpr int k2_stage_bytes(int n_tile) {
  return kK2BlockM * kK2BlockKBytes + n_tile * kK2BlockKBytes;
}
// epilogue tables (svnorm + pair weights) and the half-exchange buffer live in smem too
__host__ __device__ constexpr int k2_table_bytes(int n_pad, int n_pairs) {
  return ((n_pad * (4 + 8 * n_pairs) + 127) / 128) * 128 + kK2BlockM * n_pairs * 8;
}
__host__ __device__ constexpr int k2_smem_bytes(int n_tile, int stages, int n_pad, int n_pairs) {
  return stages * k2_stage_bytes(n_tile) + 1024 /*align slack*/ + 256 /*barriers*/ +
         k2_table_bytes(n_pad, n_pairs);
}
inline int k2_pick_stages(int n_tile, int n_pad, int n_pairs) {
  int s = (232448 - 1024 - 256 - k2_table_bytes(n_pad, n_pairs)) / k2_stage_bytes(n_tile);
  return s > kK2MaxStages ? kK2MaxStages : s;
}

// exp(x) for x <= 0 to ~5e-9 relative: exact range reduction in fp64, the r^3.. tail of the
// series in fp32, 1 + r + r^2/2 + tail reassembled in fp64 (the epilogue needs ~1e-8; the full
// fp64 exp() was the latency bottleneck of this kernel).
__device__ __forceinline__ double exp_neg_fast(double x) {
  const double fn = rint(x * 1.4426950408889634074);
  double r = fma(fn, -6.93147180369123816490e-01, x);
  r = fma(fn, -1.90821492927058770002e-10, r);
  const float rf = static_cast<float>(r);
  float q = fmaf(rf, 2.7557319e-6f, 2.4801587e-5f);   // 1/9!, 1/8!
  q = fmaf(q, rf, 1.9841270e-4f);                      // 1/7!
  q = fmaf(q, rf, 1.3888889e-3f);                      // 1/6!
  q = fmaf(q, rf, 8.3333333e-3f);                      // 1/5!
  q = fmaf(q, rf, 4.1666667e-2f);                      // 1/4!
  q = fmaf(q, rf, 1.6666667e-1f);                      // 1/3!
  const float tail = q * rf * rf * rf;
  const double e = fma(r, fma(r, 0.5, 1.0), 1.0) + static_cast<double>(tail);
  const int n = static_cast<int>(fn);
  if (n < -1000) return 0.0;
  return __hiloint2double(__double2hiint(e) + (n << 20), __double2loint(e));
}

// scipy.special.expit in float64
__device__ __forceinline__ double expit_f64(double x) {
  if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
  const double e = exp(x);
  return e / (1.0 + e);
}

// The O(C^2) tail shared by the RBF and linear scorers: decision values f[] (already OvR
// for C > 2) -> Platt -> normalise -> clip -> argmax -> threshold.
template <int C>
__device__ __forceinline__ void platt_argmax_store(const double (&f)[C], const double* platt_a,
                                                   const double* platt_b, double min_proba,
                                                   int64_t b, float* proba, float* decision,
                                                   int32_t* label, uint8_t* known) {
  double P[C];
  if (C == 2) {
    P[1] = expit_f64(-(platt_a[0] * f[0] + platt_b[0]));
    P[0] = 1.0 - P[1];
    if (decision) decision[b] = static_cast<float>(f[0]);
  } else {
    double den = 0.0;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      P[k] = expit_f64(-(platt_a[k] * f[k] + platt_b[k]));
      den += P[k];
      if (decision) decision[b * C + k] = static_cast<float>(f[k]);
    }
#pragma unroll
    for (int k = 0; k < C; ++k) P[k] = (den != 0.0) ? P[k] / den : 1.0 / C;
  }
  int best = 0;
  double pbest = 0.0;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    if (1.0 < P[k] && P[k] <= 1.0 + 1e-5) P[k] = 1.0;
    if (k == 0 || P[k] > pbest) {   // np.argmax: first maximum wins
      best = k;
      pbest = P[k];
    }
  }
#pragma unroll
  for (int k = 0; k < C; ++k) proba[b * C + k] = static_cast<float>(P[k]);
  label[b] = best;
  if (known) known[b] = (pbest >= min_proba) ? 1 : 0;
}

// OvO decision values (libsvm pair order) -> what SVC.decision_function returns.
template <int C>
__device__ __forceinline__ void ovr_transform(const double (&dec)[C * (C - 1) / 2], double (&f)[C]) {
  if constexpr (C == 2) {
    f[0] = -dec[0];  // SK/svm/_base.py binary sign flip
    f[1] = 0.0;
    return;
  } else {
  double votes[C], soc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) votes[k] = soc[k] = 0.0;
  int q = 0;
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int j = i + 1; j < C; ++j) {
      const double conf = -dec[q];
      soc[i] -= conf;
      soc[j] += conf;
      if (dec[q] < 0.0) votes[j] += 1.0; else votes[i] += 1.0;
      ++q;
    }
#pragma unroll
  for (int k = 0; k < C; ++k) f[k] = votes[k] + soc[k] / (3.0 * (fabs(soc[k]) + 1.0));
  }
}

template <int C>
__global__ void __launch_bounds__(kK2Threads, 1)
k2_rbf_i8(const __grid_constant__ CUtensorMap map_feats, const __grid_constant__ CUtensorMap map_sv,
          const K2Params p) {
  constexpr int NP = C * (C - 1) / 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int stage_bytes = k2_stage_bytes(p.n_tile);
  const int kK2Stages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kK2Stages * stage_bytes);
  uint64_t* full = bars;                    // [stages]
  uint64_t* empty = bars + kK2MaxStages;    // [stages]
  uint64_t* tfull = empty + kK2MaxStages;   // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  unsigned char* tab = smem + kK2Stages * stage_bytes + 256;
  int32_t* s_norm = reinterpret_cast<int32_t*>(tab);                                   // [n_pad]
  double* s_w = reinterpret_cast<double*>(tab + ((p.n_pad * 4 + 7) / 8) * 8);            // [NP][n_pad]
  double* s_x = reinterpret_cast<double*>(tab + ((p.n_pad * (4 + 8 * NP) + 127) / 128) * 128);  // [128][NP]
  for (int e = threadIdx.x; e < p.n_pad; e += blockDim.x) s_norm[e] = p.svnorm[e];
  for (int e = threadIdx.x; e < p.n_pad * NP; e += blockDim.x) s_w[e] = p.pairw[e];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.B + kK2BlockM - 1) / kK2BlockM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kK2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kK2EpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_feats);
    tma_prefetch_desc(&map_sv);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (whole warp in uniform control flow, one elected lane issues: see elect_one())
    // the feature tile is re-read once per support-vector chunk: keep it in L2 until then
    const uint64_t pol_a = p.n_chunks > 1 ? policy_evict_last() : policy_evict_first();
    const uint64_t pol_b = policy_evict_last();
    uint32_t kit = 0, rs = 0, rph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (p.tile_ready) {
        // fused pipeline: the projection kernel is still running on the other SMs; wait
        // until every scan of this tile has its feature row in global memory
        const int64_t left = p.B - tile * kK2BlockM;
        const unsigned int need = left < kK2BlockM ? static_cast<unsigned int>(left) : kK2BlockM;
        while (ld_acquire_gpu(&p.tile_ready[tile]) < need) __nanosleep(200);
        fence_proxy_async_all();
      }
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        for (int kb = 0; kb < p.k_blocks; ++kb, ++kit) {
          const int s = static_cast<int>(rs);           // ring slot and its phase, carried (no runtime division)
          mbar_wait(&empty[s], rph ^ 1u);
          if (++rs == static_cast<uint32_t>(kK2Stages)) { rs = 0; rph ^= 1u; }
          if (elect_one()) {
            unsigned char* a_dst = smem + s * stage_bytes;
            unsigned char* b_dst = a_dst + kK2BlockM * kK2BlockKBytes;
            mbar_arrive_expect_tx(&full[s], stage_bytes);
            tma_load_2d(a_dst, &map_feats, kb * kK2BlockKBytes, static_cast<int32_t>(tile * kK2BlockM),
                        &full[s], pol_a);
            tma_load_2d(b_dst, &map_sv, kb * kK2BlockKBytes, ch * p.n_tile, &full[s], pol_b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc(kCS32, kFmtU8, kFmtU8, kK2BlockM, p.n_tile);
    uint32_t kit = 0, ait = 0, rs = 0, rph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int ch = 0; ch < p.n_chunks; ++ch, ++ait) {
        const int ab = ait & 1;
        mbar_wait(&tempty[ab], ((ait >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kK2MaxTileN;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++kit) {
          const int s = static_cast<int>(rs);
          mbar_wait(&full[s], rph);
          if (++rs == static_cast<uint32_t>(kK2Stages)) { rs = 0; rph ^= 1u; }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
            const uint32_t b_addr = a_addr + kK2BlockM * kK2BlockKBytes;
            const uint64_t da = umma_desc_k_sw128(a_addr);
            const uint64_t db = umma_desc_k_sw128(b_addr);
#pragma unroll
            for (int ks = 0; ks < kK2BlockKBytes / 32; ++ks) {
              // advance 32 bytes (UMMA_K = 32 for 8-bit operands) inside the swizzle span
              umma_i8(d_tmem, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc, (kb | ks) != 0);
            }
            umma_commit(&empty[s]);   // smem stage reusable once these MMAs retire
            if (kb == p.k_blocks - 1) umma_commit(&tfull[ab]);    // accumulator chunk complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // warp w serves TMEM lane quarter (w & 3); the two warps of a quarter take alternate
    // 16-column groups and the odd half hands its partial pair sums over through smem.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;       // row of the tile = scan
    uint32_t ait = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t b = tile * kK2BlockM + m;
      if (p.tile_ready) {
        const int64_t left = p.B - tile * kK2BlockM;
        const unsigned int need = left < kK2BlockM ? static_cast<unsigned int>(left) : kK2BlockM;
        while (ld_acquire_gpu(&p.tile_ready[tile]) < need) __nanosleep(200);
      }
      const int un = (b < p.B) ? p.unorm[b] : 0;
      double dec[NP];
#pragma unroll
      for (int r = 0; r < NP; ++r) dec[r] = 0.0;
      for (int ch = 0; ch < p.n_chunks; ++ch, ++ait) {
        const int ab = ait & 1;
        mbar_wait(&tfull[ab], (ait >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ab * kK2MaxTileN + (static_cast<uint32_t>(q * 32) << 16);
        for (int c0 = half * 16; c0 < p.n_tile; c0 += 32) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + c0, v);
          tmem_ld_wait();
          const int n0 = ch * p.n_tile + c0;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e;   // padded columns have zero weights (and zero SV rows)
            const int d2i = un + s_norm[n] - 2 * static_cast<int>(v[e]);
            const double kv = exp_neg_fast(static_cast<double>(d2i) * p.neg_gamma_s2);
#pragma unroll
            for (int r = 0; r < NP; ++r) dec[r] = fma(s_w[r * p.n_pad + n], kv, dec[r]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
      }
      if (half == 1) {
#pragma unroll
        for (int r = 0; r < NP; ++r) s_x[m * NP + r] = dec[r];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kK2EpiWarps * 32) : "memory");
      if (half == 0 && b < p.B) {
#pragma unroll
        for (int r = 0; r < NP; ++r) dec[r] = dec[r] + s_x[m * NP + r] - p.rho[r];
        double f[C];
        ovr_transform<C>(dec, f);
        platt_argmax_store<C>(f, p.platt_a, p.platt_b, p.min_proba, b, p.proba, p.decision,
                              p.label, p.known);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kK2EpiWarps * 32) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// General-precision scorer (float32 features, arbitrary real support vectors): direct
// sum (x - sv)^2 like libsvm, fp32 differences accumulated in fp64 per thread.  CUDA cores
// only - used when the model or the input is not integral (augmented SVs, zoomed inputs);
// the caller is told which path ran (rml_model_is_integral), it is never selected silently.
struct K2GenParams {
  int64_t B;
  int F, n_sv;
  const float* feats;      // [B][F]
  const double* sv;        // [n_sv][F] float64 exactly as SVC.support_vectors_ holds them
  const double* coef;
  const double* rho;
  const double* platt_a;
  const double* platt_b;
  double neg_gamma;
  double min_proba;
  float* proba;
  float* decision;
  int32_t* label;
  uint8_t* known;
  int class_end[kMaxClasses];
};

// grid: one CTA per 8 scans; 256 threads; each warp owns one scan, lanes stride features.
template <int C>
__global__ void __launch_bounds__(256) k2_rbf_general(const K2GenParams p) {
  constexpr int NP = C * (C - 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * 8ll + warp;
  if (b >= p.B) return;
  const float* x = p.feats + b * p.F;
  double acc[C - 1];
  double dec[NP];
#pragma unroll
  for (int r = 0; r < C - 1; ++r) acc[r] = 0.0;
#pragma unroll
  for (int r = 0; r < NP; ++r) dec[r] = 0.0;
  int cls = 0, cls_end = p.class_end[0];
  for (int n = 0; n < p.n_sv; ++n) {
    while (n >= cls_end) {
#pragma unroll
      for (int cc = 0; cc < C; ++cc)
        if (cls == cc) {
#pragma unroll
          for (int r = 0; r < C - 1; ++r) {
            const int o = (r < cc) ? r : r + 1;
            const int lo = (cc < o) ? cc : o, hi = (cc < o) ? o : cc;
            dec[lo * (2 * C - lo - 1) / 2 + (hi - lo - 1)] += acc[r];
            acc[r] = 0.0;
          }
        }
      ++cls;
      cls_end = p.class_end[cls];
    }
    const double* s = p.sv + static_cast<int64_t>(n) * p.F;
    double d2 = 0.0;
    for (int f0 = lane; f0 < p.F; f0 += 32) {
      const double d = static_cast<double>(x[f0]) - s[f0];   // X cast to f64, SK/svm/_base.py:590
      d2 = fma(d, d, d2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    const double kv = exp(p.neg_gamma * d2);
#pragma unroll
    for (int r = 0; r < C - 1; ++r) acc[r] = fma(p.coef[r * p.n_sv + n], kv, acc[r]);
  }
#pragma unroll
  for (int cc = 0; cc < C; ++cc)
    if (cls == cc) {
#pragma unroll
      for (int r = 0; r < C - 1; ++r) {
        const int o = (r < cc) ? r : r + 1;
        const int lo = (cc < o) ? cc : o, hi = (cc < o) ? o : cc;
        dec[lo * (2 * C - lo - 1) / 2 + (hi - lo - 1)] += acc[r];
      }
    }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < NP; ++r) dec[r] -= p.rho[r];
    double f[C];
    ovr_transform<C>(dec, f);
    platt_argmax_store<C>(f, p.platt_a, p.platt_b, p.min_proba, b, p.proba, p.decision, p.label,
                          p.known);
  }
}

// ------------------------------------------------------------------------------------------
// Linear scorer: SGDClassifier(loss='log').decision_function = X coef^T + intercept
// (train.py:368-369; the model predict.log:9 actually deployed), then the same Platt tail.
// HBM-bound GEMV-like op: one warp per scan, coef (C x F doubles) stays in L1/L2.
struct K2LinParams {
  int64_t B;
  int F;
  int stride;              // elements per feature row
  int dtype;               // 0 f32 (already scaled), 1 u8 raw (scale folded: value / feature_scale)
  const void* feats;
  const double* coef;      // [Crows][F]
  const double* intercept; // [Crows]
  const double* platt_a;
  const double* platt_b;
  double inv_scale;        // 1 / feature_scale for u8 input
  double feature_scale;
  double min_proba;
  float* proba;
  float* decision;
  int32_t* label;
  uint8_t* known;
};

template <int C>
__global__ void __launch_bounds__(256) k2_linear(const K2LinParams p) {
  constexpr int R = (C == 2) ? 1 : C;   // sklearn keeps one row for binary problems
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * 8ll + warp;
  if (b >= p.B) return;
  double s[R];
#pragma unroll
  for (int r = 0; r < R; ++r) s[r] = 0.0;
  if (p.dtype == 0) {
    const float* x = reinterpret_cast<const float*>(p.feats) + b * p.stride;
    for (int f0 = lane; f0 < p.F; f0 += 32) {
      const double xv = static_cast<double>(x[f0]);
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = fma(xv, __ldg(&p.coef[r * p.F + f0]), s[r]);
    }
  } else {
    const uint8_t* x = reinterpret_cast<const uint8_t*>(p.feats) + b * p.stride;
    const float fs = static_cast<float>(p.feature_scale);
    for (int f0 = lane; f0 < p.F; f0 += 32) {
      // reproduce the reference's float32 feature u/255 (common.py:148) before the f64 dot
      const double xv = static_cast<double>(__fdiv_rn(static_cast<float>(x[f0]), fs));
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = fma(xv, __ldg(&p.coef[r * p.F + f0]), s[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
  }
  if (lane == 0) {
    double f[C];
#pragma unroll
    for (int r = 0; r < C; ++r) f[r] = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = s[r] + p.intercept[r];
    platt_argmax_store<C>(f, p.platt_a, p.platt_b, p.min_proba, b, p.proba, p.decision, p.label,
                          p.known);
  }
}

}  // namespace rml
