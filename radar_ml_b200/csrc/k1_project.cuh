// K1 project_scale — replaces predict.py:102-107 (projection extraction) and
// common.py:141-149 (concat xz|yz|xy, optional /RADAR_MAX) for batches of radar cubes.
//
// Fast path (arena 22x31x176, common.py:25-27): one persistent CTA per SM streams the cube
// as 22 contiguous i-slabs (31x176 fp32 = 21 824 B) through a ring of shared-memory stages
// filled by 1-D bulk async copies (UBLKCP) that complete on mbarriers.  Warp roles:
//   warps 0..7  "row" warps: own 4 rows j of every slab; running max over i (yz) lives in
//               registers, the per-row max over k (xy) is one CREDUX.MAX.F32 per row
//   warps 8..11 "column" warps: every 4th slab each, max over j for every k (xz)
//   warp 12     producer: issues the bulk copies (one elected lane)
//   warp 13     flusher: writes the finished u8 feature row back (bulk S2G); float32 rows are stored
//               straight from the row / column warps (coalesced float2 segments), no staging
// One HBM read of the cube, no re-reads; output is 2 % (u8) / 8 % (f32) of the input bytes.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace rml {

constexpr int kSX = 22, kSY = 31, kSZ = 176;
constexpr int kSlabElems = kSY * kSZ;              // 5456
constexpr int kSlabBytes = kSlabElems * 4;         // 21824
constexpr int kCubeElems = kSX * kSlabElems;       // 120032
constexpr int kFxz = kSX * kSZ, kFyz = kSY * kSZ, kFxy = kSX * kSY;  // 3872, 5456, 682
constexpr int kRowWarps = 8, kColWarps = 4;
constexpr int kK1Threads = (kRowWarps + kColWarps + 2) * 32;  // 448
// ring depth: 8 slabs (175 KB in flight per SM).  u8 rows are staged in shared memory (one bulk store
// per scan, the co-resident scorer is told when it has landed); float32 rows go STRAIGHT to global
// memory from the warps that produce them — every store is a coalesced float2 row segment — so they
// need no staging and the ring keeps its full depth (with two 40 KB staging rows it had 4 stages and
// the kernel was latency-bound at 6.1 TB/s).
template <typename OutT>
struct K1Cfg {
  static constexpr int kStages = 8;
  static constexpr bool kDirect = sizeof(OutT) == 4;
};
// A column warp only waits on the `full` barriers of the slabs it owns, so it must also own the
// previous use of that ring slot (mbarrier parity waits cannot tell phase n from phase n-2).
static_assert(K1Cfg<uint8_t>::kStages % kColWarps == 0 && K1Cfg<float>::kStages % kColWarps == 0,
              "ring depth must be a multiple of the column-warp count");

struct K1Params {
  const float* cubes;
  void* feats;        // [B][stride] u8 or f32
  int32_t* norms;     // [B] (u8 only, nullable)
  unsigned int* status;  // [0] += number of non-integral values seen (u8 only)
  int64_t B;
  int stride;         // elements per output row
  int F;              // valid features per row
  uint32_t mask;
  float offset, scale;  // f32 output: (v - offset) / scale when affine != 0
  int affine;
  const float* aff_off;  // nullable: per-feature offset[F] / scale[F] tables (rml_load_affine, a fitted
  const float* aff_scl;  // StandardScaler's mean_ / scale_); they replace the scalar pair
  int split;                // bulk copies per slab (1; 2 / 4 are tuning experiments)
  unsigned int* tile_done;  // nullable: [ceil(B/128)] += 1 per finished scan (u8 path) so a
                            // co-resident scorer can start on a 128-scan tile as soon as it is whole
  // SUMS variant (common.py:45-80 DerivedTarget.get_derived_targets in the SAME pass over the cube):
  int32_t* dt_ijk;          // [B][dt_T][3], ascending by axis sum like np.argsort (last = strongest)
  float* dt_sums;           // nullable: [B][22 + 31 + 176] theta | phi | r axis sums
  int dt_T;
};

constexpr int kK1MaxTargets = 8;
// shared memory of the SUMS variant, in floats, per staging parity: per-slab theta partials of the 8 row
// warps | phi sums | per-warp r partials; plus one final theta | phi | r array for the ranking
constexpr int kK1SumXs = kSX * kRowWarps, kK1SumSy = 32, kK1SumZp = kRowWarps * kSZ;
constexpr int kK1SumBuf = kK1SumXs + kK1SumSy + kK1SumZp;
constexpr int kK1SumFloats = 2 * kK1SumBuf + 232;

// np.max propagates NaN; fmaxf drops it.  max.NaN.f32 (FMNMX.NAN) costs the same instruction.
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float warp_max_nan_f32(float v) {
  float r;
  asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// (v - offset) / scale as an IEEE float32 true division (common.py:148: x/255f != x*(1/255f) for
// 126 of the 256 integers); per-feature tables when loaded
__device__ __forceinline__ float affine_apply(float v, float offset, float scale, const float* off,
                                              const float* scl, int idx) {
  return off ? __fdiv_rn(v - __ldg(off + idx), __ldg(scl + idx)) : __fdiv_rn(v - offset, scale);
}

template <typename OutT>
struct Emit;

template <>
struct Emit<uint8_t> {
  // raw integer value; flags anything that is not an integer in [0,255]
  static __device__ __forceinline__ uint32_t cvt(float v, uint32_t& bad) {
    uint32_t u = __float2uint_rn(v);
    bad |= (static_cast<float>(u) != v) | (u > 255u);
    return u & 255u;
  }
  static __device__ __forceinline__ void put2(uint8_t* stg, int idx, float a, float b,
                                              const K1Params&, uint32_t& sumsq, uint32_t& bad) {
    uint32_t ua = cvt(a, bad), ub = cvt(b, bad);
    sumsq += ua * ua + ub * ub;
    *reinterpret_cast<uint16_t*>(stg + idx) = static_cast<uint16_t>(ua | (ub << 8));
  }
  static __device__ __forceinline__ void put1(uint8_t* stg, int idx, float a, const K1Params&,
                                              uint32_t& sumsq, uint32_t& bad) {
    uint32_t ua = cvt(a, bad);
    sumsq += ua * ua;
    stg[idx] = static_cast<uint8_t>(ua);
  }
};

template <>
struct Emit<float> {
  static __device__ __forceinline__ float cvt(float v, const K1Params& p, int idx) {
    return p.affine ? affine_apply(v, p.offset, p.scale, p.aff_off, p.aff_scl, idx) : v;
  }
  static __device__ __forceinline__ void put2(float* stg, int idx, float a, float b,
                                              const K1Params& p, uint32_t&, uint32_t&) {
    *reinterpret_cast<float2*>(stg + idx) = make_float2(cvt(a, p, idx), cvt(b, p, idx + 1));
  }
  static __device__ __forceinline__ void put1(float* stg, int idx, float a, const K1Params& p,
                                              uint32_t&, uint32_t&) {
    stg[idx] = cvt(a, p, idx);
  }
};

template <typename OutT>
__host__ __device__ constexpr int k1_staging_bytes() {
  // u8: K-padded row (10112 B for the full mask); f32: 10010*4 rounded up to 128
  return sizeof(OutT) == 1 ? 10112 : 40064;
}
template <typename OutT, bool SUMS = false>
__host__ __device__ constexpr int k1_smem_bytes() {
  return K1Cfg<OutT>::kStages * kSlabBytes + (K1Cfg<OutT>::kDirect ? 0 : 2 * k1_staging_bytes<OutT>()) + 256 +
         (SUMS ? kK1SumFloats * 4 : 0);
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Row-warp role of k1_project_max: NR rows j of every slab (NR compile-time so the 12 smem
// loads of a slab are issued back to back and the NR warp reductions overlap).
template <typename OutT, int NR, bool SUMS>
__device__ __forceinline__ void k1_row_warp(const K1Params& p, const float* slabs, OutT* stg0,
                                            uint64_t* full, uint64_t* empty, uint64_t* done,
                                            uint64_t* sfree, uint32_t* norm_acc, float* sums_s, int warp, int lane) {
  constexpr bool kSync = !K1Cfg<OutT>::kDirect || SUMS;     // the flusher has work per scan
  constexpr int kStgBytes = k1_staging_bytes<OutT>();
  constexpr int kK1Stages = K1Cfg<OutT>::kStages;
  const float NEG = -FLT_MAX;
  const int off_yz = (p.mask & 1u) ? kFxz : 0;
  const int off_xy = off_yz + ((p.mask & 2u) ? kFyz : 0);
  const int j0 = warp * 4;
  float yz[NR][6];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < 6; ++m) yz[r][m] = NEG;
  uint32_t sumsq = 0, bad = 0;
  uint32_t it = 0, t = 0;
  const bool third = lane < 24;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++t) {
    const int buf = t & 1;
    OutT* stg;
    if (K1Cfg<OutT>::kDirect) stg = reinterpret_cast<OutT*>(p.feats) + b * static_cast<int64_t>(p.stride);   // the global row itself
    else stg = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(stg0) + buf * kStgBytes);
    if (kSync) mbar_wait(&sfree[buf], ((t >> 1) & 1) ^ 1);
    float* sb = sums_s + buf * kK1SumBuf;           // SUMS: this parity's partial-sum area
    float ysum[NR], zs[6];
    if (SUMS) {
#pragma unroll
      for (int r = 0; r < NR; ++r) ysum[r] = 0.f;
#pragma unroll
      for (int q = 0; q < 6; ++q) zs[q] = 0.f;
    }
    for (int i = 0; i < kSX; ++i, ++it) {
      const int stage = it % kK1Stages;
      mbar_wait(&full[stage], (it / kK1Stages) & 1);
      const float2* rows = reinterpret_cast<const float2*>(slabs + stage * kSlabElems + j0 * kSZ);
      float2 a[NR], c[NR], e[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        a[r] = rows[r * (kSZ / 2) + lane];
        c[r] = rows[r * (kSZ / 2) + 32 + lane];
        e[r] = third ? rows[r * (kSZ / 2) + 64 + lane] : make_float2(NEG, NEG);
      }
      float m[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        yz[r][0] = max_nan(yz[r][0], a[r].x);
        yz[r][1] = max_nan(yz[r][1], a[r].y);
        yz[r][2] = max_nan(yz[r][2], c[r].x);
        yz[r][3] = max_nan(yz[r][3], c[r].y);
        yz[r][4] = max_nan(yz[r][4], e[r].x);
        yz[r][5] = max_nan(yz[r][5], e[r].y);
        m[r] = max_nan(max_nan(max_nan(a[r].x, a[r].y), max_nan(c[r].x, c[r].y)), max_nan(e[r].x, e[r].y));
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) m[r] = warp_max_nan_f32(m[r]);
      if (SUMS) {
        // axis sums of the raw cube from the values already in registers: phi (per row j, over i and
        // k) and r (per k, over i and j) stay per-lane partials until the end of the scan; theta (per
        // slab i) is reduced here and left for the flusher, one partial per row warp
        float xs = 0.f;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const float ex = third ? e[r].x : 0.f, ey = third ? e[r].y : 0.f;
          const float s6 = ((a[r].x + a[r].y) + (c[r].x + c[r].y)) + (ex + ey);
          ysum[r] += s6;
          xs += s6;
          zs[0] += a[r].x; zs[1] += a[r].y; zs[2] += c[r].x; zs[3] += c[r].y; zs[4] += ex; zs[5] += ey;
        }
        xs = warp_sum_f32(xs);
        if (lane == 0) sb[i * kRowWarps + warp] = xs;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);    // the slab has been consumed
      // lane r stores the xy value of row r: one conversion per slab instead of NR
      float mine = m[0];
#pragma unroll
      for (int r = 1; r < NR; ++r) mine = (lane == r) ? m[r] : mine;
      if (lane < NR && (p.mask & 4u))
        Emit<OutT>::put1(stg, off_xy + i * kSY + j0 + lane, mine, p, sumsq, bad);
    }
    // end of scan: the running max over i is the yz projection
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (p.mask & 2u) {
        const int base = off_yz + (j0 + r) * kSZ + 2 * lane;
        Emit<OutT>::put2(stg, base, yz[r][0], yz[r][1], p, sumsq, bad);
        Emit<OutT>::put2(stg, base + 64, yz[r][2], yz[r][3], p, sumsq, bad);
        if (third) Emit<OutT>::put2(stg, base + 128, yz[r][4], yz[r][5], p, sumsq, bad);
      }
#pragma unroll
      for (int m2 = 0; m2 < 6; ++m2) yz[r][m2] = NEG;
    }
    if (sizeof(OutT) == 1) {
      const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
      if (lane == 0 && tot) atomicAdd(&norm_acc[buf], tot);
      sumsq = 0;
    }
    if (SUMS) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const float y = warp_sum_f32(ysum[r]);
        if (lane == 0) sb[kK1SumXs + j0 + r] = y;
      }
      float* zp = sb + kK1SumXs + kK1SumSy + warp * kSZ + 2 * lane;
      *reinterpret_cast<float2*>(zp) = make_float2(zs[0], zs[1]);
      *reinterpret_cast<float2*>(zp + 64) = make_float2(zs[2], zs[3]);
      if (third) *reinterpret_cast<float2*>(zp + 128) = make_float2(zs[4], zs[5]);
    }
    if (kSync) {
      if (!K1Cfg<OutT>::kDirect) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[buf]);
    }
  }
  if (bad) atomicAdd(p.status, 1u);
}

// SUMS: the flusher turns the partial sums of one scan into DerivedTarget's output (common.py:45-80):
// theta | phi | r axis sums (fixed summation order: reproducible) and, per axis, the dt_T indices with
// the largest sums, ascending like np.argsort (same ranking rules as k0_derive_targets).
__device__ __forceinline__ void k1_rank_targets(const K1Params& p, const float* sb, float* S, int64_t b, int lane) {
  constexpr int kN = kSX + kSY + kSZ;
  for (int e = lane; e < kN; e += 32) {
    float v = 0.f;
    if (e < kSX) {
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) v += sb[e * kRowWarps + w];
    } else if (e < kSX + kSY) {
      v = sb[kK1SumXs + e - kSX];
    } else {
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) v += sb[kK1SumXs + kK1SumSy + w * kSZ + e - kSX - kSY];
    }
    S[e] = v;
    if (p.dt_sums) p.dt_sums[b * kN + e] = v;
  }
  __syncwarp();
  for (int axis = 0; axis < 3; ++axis) {
    float* A = axis == 0 ? S : (axis == 1 ? S + kSX : S + kSX + kSY);
    const int n = axis == 0 ? kSX : (axis == 1 ? kSY : kSZ);
    for (int t = 0; t < p.dt_T; ++t) {
      float best = -FLT_MAX;
      int bi = 0x7fffffff;
      for (int e = lane; e < n; e += 32) {
        const float v = A[e];
        if (v > best) { best = v; bi = e; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      __syncwarp();
      if (lane == 0) {
        if (bi >= n) {     // NaN / -inf sums: lowest unused index, reported (status[3] -> RML_E_INVALID)
          bi = 0;
          while (bi < n - 1 && A[bi] == -FLT_MAX) ++bi;
          if (p.status) atomicAdd(p.status + 3, 1u);
        }
        p.dt_ijk[(b * p.dt_T + (p.dt_T - 1 - t)) * 3 + axis] = bi;
        A[bi] = -FLT_MAX;
      }
      __syncwarp();
    }
  }
}

template <typename OutT, bool SUMS = false>
__global__ void __launch_bounds__(kK1Threads, 1) k1_project_max(const K1Params p) {
  constexpr bool kSync = !K1Cfg<OutT>::kDirect || SUMS;
  constexpr int kK1Stages = K1Cfg<OutT>::kStages;
  extern __shared__ __align__(128) unsigned char smem[];
  float* slabs = reinterpret_cast<float*>(smem);
  OutT* stg0 = reinterpret_cast<OutT*>(smem + kK1Stages * kSlabBytes);
  constexpr int kStgBytes = k1_staging_bytes<OutT>();
  unsigned char* tail = smem + kK1Stages * kSlabBytes + (K1Cfg<OutT>::kDirect ? 0 : 2 * kStgBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);   // [kK1Stages]
  uint64_t* empty = full + kK1Stages;                   // [kK1Stages]
  uint64_t* done = empty + kK1Stages;                   // [2] scan finished in staging buf
  uint64_t* sfree = done + 2;                           // [2] staging buf flushed
  uint32_t* norm_acc = reinterpret_cast<uint32_t*>(sfree + 2);  // [2]
  float* sums_s = reinterpret_cast<float*>(tail + 256);         // SUMS: [2][kK1SumBuf] partials | [232] final

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kK1Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kRowWarps + 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&done[b], kRowWarps + kColWarps);
      mbar_init(&sfree[b], 1);
      norm_acc[b] = 0;
    }
    fence_barrier_init();
  }
  // zero both staging rows once: pad bytes [F, stride) stay zero for the whole kernel
  if (!K1Cfg<OutT>::kDirect)
    for (int i = threadIdx.x; i < 2 * kStgBytes / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(stg0)[i] = 0u;
  __syncthreads();

  const int off_xz = 0;
  const float NEG = -FLT_MAX;

  uint32_t it = 0;  // running slab counter (identical in every role)
  uint32_t t = 0;   // running scan counter of this CTA

  if (warp < kRowWarps) {
    // ------------------------------------------------------------------ row warps
    if (warp == kRowWarps - 1)
      k1_row_warp<OutT, kSY - 4 * (kRowWarps - 1), SUMS>(p, slabs, stg0, full, empty, done, sfree, norm_acc, sums_s, warp, lane);
    else
      k1_row_warp<OutT, 4, SUMS>(p, slabs, stg0, full, empty, done, sfree, norm_acc, sums_s, warp, lane);
  } else if (warp < kRowWarps + kColWarps) {
    // ------------------------------------------------------------------ column warps
    const int c = warp - kRowWarps;
    uint32_t sumsq = 0, bad = 0;
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++t) {
      const int buf = t & 1;
      OutT* stg;
      if (K1Cfg<OutT>::kDirect) stg = reinterpret_cast<OutT*>(p.feats) + b * static_cast<int64_t>(p.stride);
      else stg = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(stg0) + buf * kStgBytes);
      if (kSync) mbar_wait(&sfree[buf], ((t >> 1) & 1) ^ 1);
      for (int i = 0; i < kSX; ++i, ++it) {
        if ((it % kColWarps) != static_cast<uint32_t>(c)) continue;   // ownership follows the ring slot
        const int stage = it % kK1Stages;
        mbar_wait(&full[stage], (it / kK1Stages) & 1);
        if (p.mask & 1u) {
          const float2* slab = reinterpret_cast<const float2*>(slabs + stage * kSlabElems);
          float m0 = NEG, m1 = NEG, m2 = NEG, m3 = NEG, m4 = NEG, m5 = NEG;
#pragma unroll 8
          for (int j = 0; j < kSY; ++j) {
            const float2* row = slab + j * (kSZ / 2);
            const float2 a = row[lane];
            const float2 cc = row[32 + lane];
            float2 e = make_float2(NEG, NEG);
            if (lane < 24) e = row[64 + lane];
            m0 = max_nan(m0, a.x);
            m1 = max_nan(m1, a.y);
            m2 = max_nan(m2, cc.x);
            m3 = max_nan(m3, cc.y);
            m4 = max_nan(m4, e.x);
            m5 = max_nan(m5, e.y);
          }
          const int base = off_xz + i * kSZ + 2 * lane;
          Emit<OutT>::put2(stg, base, m0, m1, p, sumsq, bad);
          Emit<OutT>::put2(stg, base + 64, m2, m3, p, sumsq, bad);
          if (lane < 24) Emit<OutT>::put2(stg, base + 128, m4, m5, p, sumsq, bad);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      if (sizeof(OutT) == 1) {
        const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
        if (lane == 0 && tot) atomicAdd(&norm_acc[buf], tot);
        sumsq = 0;
      }
      if (kSync) {
        if (!K1Cfg<OutT>::kDirect) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[buf]);
      }
    }
    if (bad) atomicAdd(p.status, 1u);
  } else if (warp == kRowWarps + kColWarps) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        const float* cube = p.cubes + b * kCubeElems;
        for (int i = 0; i < kSX; ++i, ++it) {
          const int stage = it % kK1Stages;
          mbar_wait(&empty[stage], ((it / kK1Stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[stage], kSlabBytes);
          const int part = kSlabElems / p.split;
          for (int q = 0; q < p.split; ++q)
            bulk_g2s(slabs + stage * kSlabElems + q * part, cube + i * kSlabElems + q * part, part * 4,
                     &full[stage], pol);
        }
      }
    }
  } else if (kSync) {
    // ------------------------------------------------------------------ flusher (u8 rows, axis sums)
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++t) {
      const int buf = t & 1;
      OutT* stg = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(stg0) + buf * kStgBytes);
      mbar_wait(&done[buf], (t >> 1) & 1);
      if (sizeof(OutT) == 1 && lane == 0) {
        bulk_s2g(reinterpret_cast<uint8_t*>(p.feats) + b * p.stride, stg, p.stride);
        bulk_commit();
        if (p.norms) p.norms[b] = static_cast<int32_t>(norm_acc[buf]);
        norm_acc[buf] = 0;
      }
      if (SUMS) k1_rank_targets(p, sums_s + buf * kK1SumBuf, sums_s + 2 * kK1SumBuf, b, lane);   // under the bulk store
      if (K1Cfg<OutT>::kDirect) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfree[buf]);
      } else if (sizeof(OutT) == 1) {
        if (lane == 0) {
          bulk_wait_read<0>();
          mbar_arrive(&sfree[buf]);
          if (p.tile_done) {
            bulk_wait_all<0>();          // the feature row has landed in global memory
            __threadfence();
            atomicAdd(&p.tile_done[b >> 7], 1u);
          }
        }
      }
    }
    if (sizeof(OutT) == 1 && lane == 0) bulk_wait_all<0>();
  }
}

// ------------------------------------------------------------------------------------------
// Generic kernels: any arena, MAX or SLICE (predict.py:102-107 semantics incl. numpy negative
// index wrap).  One CTA per scan; SLICE touches only the three planes it needs.
struct K1GenParams {
  const void* cubes;   // float32 or uint8 voxels (the kernels' InT)
  const int32_t* ijk;  // [B][3] (slice)
  void* feats;
  int32_t* norms;
  unsigned int* status;
  int64_t B;
  int sx, sy, sz;
  int stride, F;
  uint32_t mask;
  float offset, scale;
  int affine;
  const float* aff_off;   // nullable per-feature tables, see K1Params
  const float* aff_scl;
  int mode;
  int64_t cube_stride;    // elements between consecutive scans' cubes; 0 = every row reads cube 0
                          // (one scan, T targets: predict.py:93-119)
};

template <typename OutT>
__device__ __forceinline__ void gen_put(OutT* out, int idx, float v, const K1GenParams& p,
                                        uint32_t& sumsq, uint32_t& bad);
template <>
__device__ __forceinline__ void gen_put<uint8_t>(uint8_t* out, int idx, float v,
                                                 const K1GenParams&, uint32_t& sumsq,
                                                 uint32_t& bad) {
  uint32_t u = Emit<uint8_t>::cvt(v, bad);
  sumsq += u * u;
  out[idx] = static_cast<uint8_t>(u);
}
template <>
__device__ __forceinline__ void gen_put<float>(float* out, int idx, float v, const K1GenParams& p,
                                               uint32_t&, uint32_t&) {
  out[idx] = p.affine ? affine_apply(v, p.offset, p.scale, p.aff_off, p.aff_scl, idx) : v;
}

template <typename OutT, typename InT = float>
__global__ void __launch_bounds__(256) k1_project_generic(const K1GenParams p) {
  __shared__ uint32_t s_sum;
  const int sx = p.sx, sy = p.sy, sz = p.sz;
  const int fxz = sx * sz, fyz = sy * sz, fxy = sx * sy;
  const int off_yz = (p.mask & 1u) ? fxz : 0;
  const int off_xy = off_yz + ((p.mask & 2u) ? fyz : 0);
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const InT* cube = static_cast<const InT*>(p.cubes) + b * p.cube_stride;
    OutT* out = reinterpret_cast<OutT*>(p.feats) + b * static_cast<int64_t>(p.stride);
    uint32_t sumsq = 0, bad = 0;
    int ti = 0, tj = 0, tk = 0;
    bool ok = true;
    if (p.mode == 1) {
      ti = p.ijk[b * 3 + 0];
      tj = p.ijk[b * 3 + 1];
      tk = p.ijk[b * 3 + 2];
      // numpy semantics: negative indices wrap once, anything else is an IndexError
      if (ti < 0) ti += sx;
      if (tj < 0) tj += sy;
      if (tk < 0) tk += sz;
      ok = ti >= 0 && ti < sx && tj >= 0 && tj < sy && tk >= 0 && tk < sz;
      if (!ok) {
        if (threadIdx.x == 0) atomicAdd(p.status + 1, 1u);
        ti = tj = tk = 0;
      }
    }
    if (p.mask & 1u) {
      for (int e = threadIdx.x; e < fxz; e += blockDim.x) {
        const int i = e / sz, k = e - i * sz;
        float v;
        if (p.mode == 1) {
          v = static_cast<float>(cube[(static_cast<int64_t>(i) * sy + tj) * sz + k]);
        } else {
          v = -FLT_MAX;
          for (int j = 0; j < sy; ++j) v = max_nan(v, static_cast<float>(cube[(static_cast<int64_t>(i) * sy + j) * sz + k]));
        }
        gen_put<OutT>(out, e, ok ? v : 0.f, p, sumsq, bad);
      }
    }
    if (p.mask & 2u) {
      for (int e = threadIdx.x; e < fyz; e += blockDim.x) {
        const int j = e / sz, k = e - j * sz;
        float v;
        if (p.mode == 1) {
          v = static_cast<float>(cube[(static_cast<int64_t>(ti) * sy + j) * sz + k]);
        } else {
          v = -FLT_MAX;
          for (int i = 0; i < sx; ++i) v = max_nan(v, static_cast<float>(cube[(static_cast<int64_t>(i) * sy + j) * sz + k]));
        }
        gen_put<OutT>(out, off_yz + e, ok ? v : 0.f, p, sumsq, bad);
      }
    }
    if (p.mask & 4u) {
      if (p.mode == 1) {
        for (int e = threadIdx.x; e < fxy; e += blockDim.x) {
          const float v = static_cast<float>(cube[static_cast<int64_t>(e) * sz + tk]);
          gen_put<OutT>(out, off_xy + e, ok ? v : 0.f, p, sumsq, bad);
        }
      } else {
        // one warp per (i,j) row: coalesced read along k, shuffle max
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int e = warp; e < fxy; e += nw) {
          const InT* row = cube + static_cast<int64_t>(e) * sz;
          float v = -FLT_MAX;
          for (int k = lane; k < sz; k += 32) v = max_nan(v, static_cast<float>(row[k]));
          v = warp_max_nan_f32(v);
          if (lane == 0) gen_put<OutT>(out, off_xy + e, v, p, sumsq, bad);
        }
      }
    }
    if (sizeof(OutT) == 1) {
      // zero the K padding so the row is fully defined
      for (int e = p.F + threadIdx.x; e < p.stride; e += blockDim.x) out[e] = 0;
      const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
      if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&s_sum, tot);
      __syncthreads();
      if (threadIdx.x == 0 && p.norms) p.norms[b] = static_cast<int32_t>(s_sum);
      if (bad) atomicAdd(p.status, 1u);
    }
    __syncthreads();
  }
}

// SLICE projections (predict.py:102-107) for arenas with size_z % 4 == 0: one CTA per scan, every
// thread keeps four independent 16-byte (xz, yz rows) or 4-byte (xy gather) loads in flight; the
// three planes are the only bytes touched (59 KB of the 480 KB cube, xy as 32-byte sectors).
template <typename OutT>
__device__ __forceinline__ void slice_put4(OutT* out, int idx, float4 v, const K1GenParams& p,
                                           uint32_t& sumsq, uint32_t& bad);
template <>
__device__ __forceinline__ void slice_put4<uint8_t>(uint8_t* out, int idx, float4 v, const K1GenParams&,
                                                    uint32_t& sumsq, uint32_t& bad) {
  const uint32_t a = Emit<uint8_t>::cvt(v.x, bad), b = Emit<uint8_t>::cvt(v.y, bad);
  const uint32_t c = Emit<uint8_t>::cvt(v.z, bad), d = Emit<uint8_t>::cvt(v.w, bad);
  sumsq += a * a + b * b + c * c + d * d;
  *reinterpret_cast<uint32_t*>(out + idx) = a | (b << 8) | (c << 16) | (d << 24);
}
template <>
__device__ __forceinline__ void slice_put4<float>(float* out, int idx, float4 v, const K1GenParams& p,
                                                  uint32_t&, uint32_t&) {
  if (p.affine) {
    v.x = affine_apply(v.x, p.offset, p.scale, p.aff_off, p.aff_scl, idx);
    v.y = affine_apply(v.y, p.offset, p.scale, p.aff_off, p.aff_scl, idx + 1);
    v.z = affine_apply(v.z, p.offset, p.scale, p.aff_off, p.aff_scl, idx + 2);
    v.w = affine_apply(v.w, p.offset, p.scale, p.aff_off, p.aff_scl, idx + 3);
  }
  out[idx] = v.x; out[idx + 1] = v.y; out[idx + 2] = v.z; out[idx + 3] = v.w;   // rows are 8-B aligned only
}

template <typename InT>
__device__ __forceinline__ float4 slice_load4(const InT* q);
template <>
__device__ __forceinline__ float4 slice_load4<float>(const float* q) {
  return *reinterpret_cast<const float4*>(q);
}
template <>
__device__ __forceinline__ float4 slice_load4<uint8_t>(const uint8_t* q) {
  const uchar4 u = *reinterpret_cast<const uchar4*>(q);
  return make_float4(static_cast<float>(u.x), static_cast<float>(u.y), static_cast<float>(u.z),
                     static_cast<float>(u.w));
}

template <typename OutT, typename InT = float>
__global__ void __launch_bounds__(256) k1_project_slice(const K1GenParams p) {
  __shared__ uint32_t s_sum;
  const int sx = p.sx, sy = p.sy, sz = p.sz, zq = sz >> 2;
  const int fxz = sx * sz, fyz = sy * sz, fxy = sx * sy;
  const int off_yz = (p.mask & 1u) ? fxz : 0;
  const int off_xy = off_yz + ((p.mask & 2u) ? fyz : 0);
  const int n_xz = (p.mask & 1u) ? sx * zq : 0;      // float4 items
  const int n_yz = (p.mask & 2u) ? sy * zq : 0;
  const int n_xy = (p.mask & 4u) ? fxy : 0;           // scalar items
  const int n_items = n_xz + n_yz + n_xy;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const InT* cube = static_cast<const InT*>(p.cubes) + b * p.cube_stride;
    OutT* out = reinterpret_cast<OutT*>(p.feats) + b * static_cast<int64_t>(p.stride);
    int ti = p.ijk[b * 3 + 0], tj = p.ijk[b * 3 + 1], tk = p.ijk[b * 3 + 2];
    if (ti < 0) ti += sx;                       // numpy: negative indices wrap once
    if (tj < 0) tj += sy;
    if (tk < 0) tk += sz;
    const bool ok = ti >= 0 && ti < sx && tj >= 0 && tj < sy && tk >= 0 && tk < sz;
    if (!ok) {
      if (threadIdx.x == 0) atomicAdd(p.status + 1, 1u);   // numpy would raise IndexError
      ti = tj = tk = 0;
    }
    uint32_t sumsq = 0, bad = 0;
#pragma unroll 4
    for (int e = threadIdx.x; e < n_items; e += 256) {
      if (e < n_xz) {
        const int i = e / zq, c = e - i * zq;
        float4 v = slice_load4<InT>(cube + (static_cast<int64_t>(i) * sy + tj) * sz + 4 * c);
        if (!ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
        slice_put4<OutT>(out, i * sz + 4 * c, v, p, sumsq, bad);
      } else if (e < n_xz + n_yz) {
        const int r = e - n_xz;
        const int j = r / zq, c = r - j * zq;
        float4 v = slice_load4<InT>(cube + (static_cast<int64_t>(ti) * sy + j) * sz + 4 * c);
        if (!ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
        slice_put4<OutT>(out, off_yz + j * sz + 4 * c, v, p, sumsq, bad);
      } else {
        const int r = e - n_xz - n_yz;
        const float v = ok ? static_cast<float>(cube[static_cast<int64_t>(r) * sz + tk]) : 0.f;
        gen_put<OutT>(out, off_xy + r, v, p, sumsq, bad);
      }
    }
    if (sizeof(OutT) == 1) {
      for (int e = p.F + threadIdx.x; e < p.stride; e += blockDim.x) out[e] = 0;
      const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
      if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&s_sum, tot);
      __syncthreads();
      if (threadIdx.x == 0 && p.norms) p.norms[b] = static_cast<int32_t>(s_sum);
      if (bad) atomicAdd(p.status, 1u);
    }
    __syncthreads();
  }
}

// common.process_samples on already extracted projections (common.py:141-149, zoom 1.0).
struct PsParams {
  const float* proj[3];  // xz, yz, xy (nullable when masked out)
  int len[3];            // elements per sample of each projection
  int off[3];            // output offset of each projection
  float* feats;
  int64_t B;
  int F;
  int scale;
  float offset, scale_value;   // (v - offset) / scale_value when scale != 0
  const float* aff_off;        // nullable per-feature tables, see K1Params
  const float* aff_scl;
};
__global__ void __launch_bounds__(256) k1_process_samples(const PsParams p) {
  const int64_t total = p.B * static_cast<int64_t>(p.F);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = e / p.F;
    const int f = static_cast<int>(e - b * p.F);
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (p.proj[q] && f >= p.off[q] && f < p.off[q] + p.len[q])
        v = p.proj[q][b * p.len[q] + (f - p.off[q])];
    p.feats[e] = p.scale ? affine_apply(v, p.offset, p.scale_value, p.aff_off, p.aff_scl, f) : v;
  }
}

// (n,F) float32 features scaled like common.process_samples(scale=True) -> raw u8 operand rows
// + integer norms for the tensor-core scorer.  A value qualifies when it is exactly
// float32(u)/float32(scale) for an integer u in [0,255]; anything else bumps status[0].
struct QuantParams {
  const float* feats;
  uint8_t* out;
  int32_t* norms;
  unsigned int* status;
  int64_t B;
  int F, stride;
  float scale;
};
__global__ void __launch_bounds__(256) k1_quantize(const QuantParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * 8ll + warp;
  if (b >= p.B) return;
  const float* x = p.feats + b * p.F;
  uint8_t* o = p.out + b * p.stride;
  uint32_t sumsq = 0, bad = 0;
  for (int f = lane; f < p.stride; f += 32) {
    uint32_t u = 0;
    if (f < p.F) {
      const float v = x[f];
      u = __float2uint_rn(v * p.scale);
      bad |= (u > 255u) | (__fdiv_rn(static_cast<float>(u), p.scale) != v);
      u &= 255u;
    }
    o[f] = static_cast<uint8_t>(u);
    sumsq += u * u;
  }
  const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
  if (lane == 0) p.norms[b] = static_cast<int32_t>(tot);
  if (bad) atomicAdd(p.status, 1u);
}

// common.calculate_matrix_indices (common.py:106-121 + 93-97), float64 like numpy.
struct IdxParams {
  const double* xyz;
  int32_t* ijk;
  int64_t B;
  int sx, sy, sz;
  double r_min, r_max, th_min, th_max, ph_min, ph_max;
};
__global__ void __launch_bounds__(128) k1_matrix_indices(const IdxParams p) {
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b >= p.B) return;
  const double x = p.xyz[b * 3 + 0], y = p.xyz[b * 3 + 1], z = p.xyz[b * 3 + 2];
  const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
  const double k180_pi = 180.0 / 3.141592653589793238462643383279502884;
  const double phi = __dmul_rn(atan2(y, z), k180_pi);
  const double theta = __dmul_rn(asin(x / r), k180_pi);
  // int() truncates toward zero; evaluation order as written in common.py:118-120
  const double fi = __dmul_rn(theta - p.th_min, static_cast<double>(p.sx - 1)) / (p.th_max - p.th_min);
  const double fj = __dmul_rn(phi - p.ph_min, static_cast<double>(p.sy - 1)) / (p.ph_max - p.ph_min);
  const double fk = __dmul_rn(r - p.r_min, static_cast<double>(p.sz - 1)) / (p.r_max - p.r_min);
  p.ijk[b * 3 + 0] = static_cast<int32_t>(fi);
  p.ijk[b * 3 + 1] = static_cast<int32_t>(fj);
  p.ijk[b * 3 + 2] = static_cast<int32_t>(fk);
}

}  // namespace rml
