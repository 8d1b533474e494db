// K2 general-precision scorer on the tensor cores: exact multi-digit integer distances.
//
// Non-integral features / support vectors (augmented training data train.py:84-185, zoomed
// inputs common.py:143) cannot use the single u8 x u8 product of k2_rbf_i8, and a floating
// GEMM form ||x||^2 + ||s||^2 - 2 x.s loses the distance to cancellation.  Instead every value
// v = 255 x in [0, 256) is taken to 24-bit fixed point X = round(v * 2^16) = 65536 d2 + 256 d1 + d0
// and the dot product is assembled from nine EXACT u8 x u8 -> s32 digit products:
//     X . S = sum_p 256^p  sum_{a+b=p} (A_a . B_b),           p = 0..4
// Each power p has its own s32 accumulator in TMEM (<= 3 x 6.5e8 < 2^31), the epilogue rebuilds
// the 64-bit integer ||X - S||^2 = ||X||^2 + ||S||^2 - 2 X.S exactly and hands
// exp(-gamma d^2) to the same fp64 tail as the integer kernel.  Quantisation is 2^-17 of one
// sensor count (3e-8 of a [0,1] feature); everything after it is integer arithmetic, so the
// result is deterministic and independent of the summation order.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "k2_score.cuh"
#include "ptx.cuh"

namespace rml {

constexpr int kDgTileN = 96;          // support vectors per chunk: 5 accumulators x 96 = 480 TMEM columns
constexpr int kDgStages = 2;
constexpr int kDgABytes = kK2BlockM * kK2BlockKBytes;       // 16 384 per digit
constexpr int kDgBBytes = kDgTileN * kK2BlockKBytes;        // 12 288 per digit
constexpr int kDgStageBytes = 3 * (kDgABytes + kDgBBytes);  // 86 016

struct K2DgParams {
  int64_t B;
  int n_sv, n_chunks, k_blocks, n_pad;
  const long long* unorm;      // [B]     sum X^2
  const long long* svnorm;     // [n_pad] sum S^2
  const double* pairw;         // [NP][n_pad]
  const double* rho;
  const double* platt_a;
  const double* platt_b;
  const unsigned int* tile_ready;  // nullable: [tiles] scans finished by the co-resident projection kernel
  double neg_gamma_fixed;      // -gamma / (feature_scale^2 * 2^32)
  double min_proba;
  float* proba;
  float* decision;
  int32_t* label;
  uint8_t* known;
};

struct DgMaps {
  CUtensorMap a[3];            // feature digit planes [B][kpad] u8
  CUtensorMap b[3];            // support-vector digit planes [n_pad][kpad] u8
};

__host__ __device__ constexpr int k2dg_table_bytes(int n_pad, int n_pairs) {
  return ((n_pad * (8 + 8 * n_pairs) + 127) / 128) * 128 + kK2BlockM * n_pairs * 8;
}
__host__ __device__ constexpr int k2dg_smem_bytes(int n_pad, int n_pairs) {
  return kDgStages * kDgStageBytes + 1024 + 256 + k2dg_table_bytes(n_pad, n_pairs);
}

template <int C>
__global__ void __launch_bounds__(kK2Threads, 1)
k2_rbf_digits(const __grid_constant__ DgMaps maps, const K2DgParams p) {
  constexpr int NP = C * (C - 1) / 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDgStages * kDgStageBytes);
  uint64_t* full = bars;                 // [2]
  uint64_t* empty = bars + kDgStages;    // [2]
  uint64_t* tfull = empty + kDgStages;   // [1]
  uint64_t* tempty = tfull + 1;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  unsigned char* tab = smem + kDgStages * kDgStageBytes + 256;
  long long* s_norm = reinterpret_cast<long long*>(tab);                            // [n_pad]
  double* s_w = reinterpret_cast<double*>(tab + p.n_pad * 8);                        // [NP][n_pad]
  double* s_x = reinterpret_cast<double*>(tab + ((p.n_pad * (8 + 8 * NP) + 127) / 128) * 128);
  for (int e = threadIdx.x; e < p.n_pad; e += blockDim.x) s_norm[e] = p.svnorm[e];
  for (int e = threadIdx.x; e < p.n_pad * NP; e += blockDim.x) s_w[e] = p.pairw[e];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.B + kK2BlockM - 1) / kK2BlockM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDgStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, kK2EpiWarps);
    fence_barrier_init();
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      tma_prefetch_desc(&maps.a[d]);
      tma_prefetch_desc(&maps.b[d]);
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const uint64_t pol = policy_evict_last();
    uint32_t kit = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (p.tile_ready) {
        const int64_t left = p.B - tile * kK2BlockM;
        const unsigned int need = left < kK2BlockM ? static_cast<unsigned int>(left) : kK2BlockM;
        while (ld_acquire_gpu(&p.tile_ready[tile]) < need) __nanosleep(200);
        fence_proxy_async_all();
      }
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        for (int kb = 0; kb < p.k_blocks; ++kb, ++kit) {
          const int s = kit % kDgStages;
          mbar_wait(&empty[s], ((kit / kDgStages) & 1) ^ 1);
          if (elect_one()) {
            unsigned char* a_dst = smem + s * kDgStageBytes;
            unsigned char* b_dst = a_dst + 3 * kDgABytes;
            mbar_arrive_expect_tx(&full[s], kDgStageBytes);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              tma_load_2d(a_dst + d * kDgABytes, &maps.a[d], kb * kK2BlockKBytes,
                          static_cast<int32_t>(tile * kK2BlockM), &full[s], pol);
              tma_load_2d(b_dst + d * kDgBBytes, &maps.b[d], kb * kK2BlockKBytes, ch * kDgTileN, &full[s], pol);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc(kCS32, kFmtU8, kFmtU8, kK2BlockM, kDgTileN);
    const uint32_t idesc2 = umma_idesc(kCS32, kFmtU8, kFmtU8, kK2BlockM, 2 * kDgTileN);
    uint32_t kit = 0, ait = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int ch = 0; ch < p.n_chunks; ++ch, ++ait) {
        mbar_wait(tempty, (ait & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.k_blocks; ++kb, ++kit) {
          const int s = kit % kDgStages;
          mbar_wait(&full[s], (kit / kDgStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + s * kDgStageBytes);
            const uint32_t b_addr = a_addr + 3 * kDgABytes;
            // Nine digit products per K step, six instructions: the support-vector digit tiles 0 and 1 are
            // contiguous in the stage (12 groups of 8 rows each), and so are the accumulators of powers p
            // and p+1, so A_a x [B_0 | B_1] is ONE UMMA of N = 192 into [acc_a | acc_a+1] — every A tile
            // is read from shared memory twice instead of three times.  Order: the three N = 96 products
            // with B_2 first (they initialise acc_2..4), then a = 0 (initialises acc_0 | acc_1), then
            // a = 1, 2, which only ever accumulate.
            const bool init = kb == 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const uint64_t da = umma_desc_k_sw128(a_addr + a * kDgABytes);
              const uint64_t db = umma_desc_k_sw128(b_addr + 2 * kDgBBytes);
#pragma unroll
              for (int ks = 0; ks < kK2BlockKBytes / 32; ++ks)
                umma_i8(tmem_base + (a + 2) * kDgTileN, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc, !(init && ks == 0));
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const uint64_t da = umma_desc_k_sw128(a_addr + a * kDgABytes);
              const uint64_t db = umma_desc_k_sw128(b_addr);
#pragma unroll
              for (int ks = 0; ks < kK2BlockKBytes / 32; ++ks)
                umma_i8(tmem_base + a * kDgTileN, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc2,
                        !(init && a == 0 && ks == 0));
            }
            umma_commit(&empty[s]);
            if (kb == p.k_blocks - 1) umma_commit(tfull);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    uint32_t ait = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t b = tile * kK2BlockM + m;
      if (p.tile_ready) {
        const int64_t left = p.B - tile * kK2BlockM;
        const unsigned int need = left < kK2BlockM ? static_cast<unsigned int>(left) : kK2BlockM;
        while (ld_acquire_gpu(&p.tile_ready[tile]) < need) __nanosleep(200);
      }
      const long long un = (b < p.B) ? p.unorm[b] : 0;
      double dec[NP];
#pragma unroll
      for (int r = 0; r < NP; ++r) dec[r] = 0.0;
      for (int ch = 0; ch < p.n_chunks; ++ch, ++ait) {
        mbar_wait(tfull, ait & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int c0 = half * 16; c0 < kDgTileN; c0 += 32) {
          uint32_t v0[16], v1[16], v2[16], v3[16], v4[16];
          tmem_ld_32x16(taddr + 0 * kDgTileN + c0, v0);
          tmem_ld_32x16(taddr + 1 * kDgTileN + c0, v1);
          tmem_ld_32x16(taddr + 2 * kDgTileN + c0, v2);
          tmem_ld_32x16(taddr + 3 * kDgTileN + c0, v3);
          tmem_ld_32x16(taddr + 4 * kDgTileN + c0, v4);
          tmem_ld_wait();
          const int n0 = ch * kDgTileN + c0;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e;
            const unsigned long long dot =
                static_cast<unsigned long long>(v0[e]) + (static_cast<unsigned long long>(v1[e]) << 8) +
                (static_cast<unsigned long long>(v2[e]) << 16) + (static_cast<unsigned long long>(v3[e]) << 24) +
                (static_cast<unsigned long long>(v4[e]) << 32);
            const long long d2 = un + s_norm[n] - 2ll * static_cast<long long>(dot);
            const double kv = exp_neg_fast(static_cast<double>(d2) * p.neg_gamma_fixed);
#pragma unroll
            for (int r = 0; r < NP; ++r) dec[r] = fma(s_w[r * p.n_pad + n], kv, dec[r]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
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
        platt_argmax_store<C>(f, p.platt_a, p.platt_b, p.min_proba, b, p.proba, p.decision, p.label, p.known);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kK2EpiWarps * 32) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// (n,F) float32 features scaled like common.process_samples(scale=True) -> three u8 digit planes
// of X = round((x + shift_f) * scale * 2^16) and the exact 64-bit norms.  Values outside [0, 256/scale) cannot
// be represented: they are clamped and counted in status[2] (RML_E_RANGE).
struct DgQuantParams {
  const float* feats;
  uint8_t* planes;        // [3][B][stride]
  long long* norms;       // [B]
  unsigned int* status;
  int64_t B;
  int F, stride;
  double scale;
  const double* shift;    // nullable [F]: added before scaling (standardised features -> non-negative)
};
// one warp per scan, eight scans per CTA; a lane takes 4 consecutive features per step (two float2 loads —
// rows of 10 010 floats are only 8-byte aligned — and one 32-bit store per digit plane), so every access of
// a warp is a contiguous 512 / 128-byte segment.  (The first version moved single bytes per lane and took
// 1.01 ms per 32 768 scans, 20 % of the general-precision path; the pass moves 70 KB per scan.)
__global__ void __launch_bounds__(256) k1_quantize_digits(const DgQuantParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * 8ll + warp;
  if (b >= p.B) return;
  const float* x = p.feats + b * p.F;
  const int64_t plane = p.B * static_cast<int64_t>(p.stride);
  uint8_t* o = p.planes + b * p.stride;
  const double k = p.scale * 65536.0;
  unsigned long long sumsq = 0;
  uint32_t bad = 0;
  const bool even_row = ((b * p.F) & 1) == 0;      // float2 loads need an 8-byte aligned pair
  for (int f = 4 * lane; f < p.stride; f += 128) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (even_row && f + 3 < p.F) {
      const float2 a = *reinterpret_cast<const float2*>(x + f);
      const float2 c = *reinterpret_cast<const float2*>(x + f + 2);
      v[0] = a.x; v[1] = a.y; v[2] = c.x; v[3] = c.y;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (f + q < p.F) v[q] = x[f + q];
    }
    uint32_t d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t X = 0;
      if (f + q < p.F) {
        const double xs = p.shift ? static_cast<double>(v[q]) + p.shift[f + q] : static_cast<double>(v[q]);
        const double r = rint(xs * k);              // = rint(xs * scale * 65536): the power of two is exact
        bad |= !(r >= 0.0 && r < 16777216.0);
        X = static_cast<uint32_t>(fmin(fmax(r, 0.0), 16777215.0));
      }
      d0 |= (X & 255u) << (8 * q);
      d1 |= ((X >> 8) & 255u) << (8 * q);
      d2 |= (X >> 16) << (8 * q);
      sumsq += static_cast<unsigned long long>(X) * X;
    }
    *reinterpret_cast<uint32_t*>(o + f) = d0;
    *reinterpret_cast<uint32_t*>(o + plane + f) = d1;
    *reinterpret_cast<uint32_t*>(o + 2 * plane + f) = d2;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sumsq += __shfl_xor_sync(0xffffffffu, sumsq, off);
  if (lane == 0) p.norms[b] = static_cast<long long>(sumsq);
  if (bad) atomicAdd(p.status + 2, 1u);
}

}  // namespace rml
