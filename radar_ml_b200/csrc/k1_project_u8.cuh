// K1 for uint8 cubes — the same projection + concat + scale as k1_project.cuh
// (predict.py:102-107, common.py:141-149) when the caller keeps the sensor's integers.
//
// predict.py:90-91 widens the Walabot's integer voxels (0..255) to float32; a
// caller that holds them as uint8 moves 120 032 B per scan instead of 480 128 B over PCIe and
// HBM.  Same persistent one-CTA-per-SM design and barrier protocol as k1_project_max: the cube
// streams as 22 i-slabs (31x176 B = 5 456 B) through a 16-stage ring of 1-D bulk copies; the
// arithmetic is integer SIMD, so it is exact by construction.  sm_100 has no byte-wise max
// (__vmaxu4 is a 7-instruction emulation) but VIMNMX.U16x2 / VIMNMX3.U16x2 are native: every
// word is split once by two PRMTs into its even and odd bytes (two u16x2 registers) and all
// three projections reduce on those:
//   warps 0..7  "row" warps: 4 rows j each, 22 lanes x 8 bytes per row; running max over i (yz)
//               in 2 registers per row, per-row max over k (xy) by one REDUX.MAX.U32
//   warps 8..11 "column" warps: every 4th slab, max over j for 8 bytes of k per lane (xz)
//   warp 12     producer (bulk G2S), warp 13 flusher (bulk S2G of the feature row)
// u8 output rows carry sum u^2 (DP4A) for the tensor-core scorer; float32 output applies the
// IEEE true division of common.py:148 to the widened byte.
#pragma once
#include "k1_project.cuh"

namespace rml {

constexpr int kSlabBytesU8 = kSlabElems;       // 5456
constexpr int kU8Stages = 16;                  // 87 KB in flight per SM
constexpr int kU8RowLanes = kSZ / 8;           // 22 lanes x uint2 per 176-byte row
static_assert(kU8Stages % kColWarps == 0 && kU8Stages % 2 == 0 && kSX % 2 == 0, "ring depth must be a multiple of the column-warp count");
static_assert(kSlabBytesU8 % 16 == 0 && kCubeElems % 16 == 0, "bulk copies need 16-byte granules");

struct K1U8Params {
  const uint8_t* cubes;   // [B][22][31][176] u8
  void* feats;            // [B][stride] u8 or f32
  int32_t* norms;         // [B] (u8 output only, nullable)
  int64_t B;
  int stride, F;
  uint32_t mask;
  float offset, scale;    // f32 output: (v - offset) / scale when affine != 0
  int affine;
  const float* aff_off;   // nullable per-feature tables, see K1Params
  const float* aff_scl;
  unsigned int* tile_done;  // nullable, see K1Params
};

template <typename OutT>
struct EmitB;

template <>
struct EmitB<uint8_t> {
  static __device__ __forceinline__ void put8(uint8_t* stg, int idx, uint2 w, const K1U8Params&,
                                              uint32_t& sumsq) {
    *reinterpret_cast<uint2*>(stg + idx) = w;
    sumsq = __dp4a(w.x, w.x, sumsq);
    sumsq = __dp4a(w.y, w.y, sumsq);
  }
  static __device__ __forceinline__ void put1(uint8_t* stg, int idx, uint32_t v, const K1U8Params&,
                                              uint32_t& sumsq) {
    stg[idx] = static_cast<uint8_t>(v);
    sumsq += v * v;
  }
};

template <>
struct EmitB<float> {
  static __device__ __forceinline__ float cvt(uint32_t b, const K1U8Params& p, int idx) {
    const float v = static_cast<float>(b);
    return p.affine ? affine_apply(v, p.offset, p.scale, p.aff_off, p.aff_scl, idx) : v;
  }
  static __device__ __forceinline__ float4 cvt4(uint32_t w, const K1U8Params& p, int idx) {
    return make_float4(cvt(w & 255u, p, idx), cvt((w >> 8) & 255u, p, idx + 1), cvt((w >> 16) & 255u, p, idx + 2),
                       cvt(w >> 24, p, idx + 3));
  }
  static __device__ __forceinline__ void put8(float* stg, int idx, uint2 w, const K1U8Params& p,
                                              uint32_t&) {
    *reinterpret_cast<float4*>(stg + idx) = cvt4(w.x, p, idx);
    *reinterpret_cast<float4*>(stg + idx + 4) = cvt4(w.y, p, idx + 4);
  }
  static __device__ __forceinline__ void put1(float* stg, int idx, uint32_t v, const K1U8Params& p,
                                              uint32_t&) {
    stg[idx] = cvt(v, p, idx);
  }
};

template <typename OutT>
__host__ __device__ constexpr int k1u8_smem_bytes() {
  return kU8Stages * kSlabBytesU8 + 2 * k1_staging_bytes<OutT>() + 512;
}

// 8 packed bytes -> four u16x2 registers (even / odd bytes of each word), and back
struct U16x8 {
  uint32_t e0, o0, e1, o1;
};
__device__ __forceinline__ U16x8 unpack_u8x8(uint2 v) {
  U16x8 u;
  u.e0 = __byte_perm(v.x, 0u, 0x4240);   // bytes 0,2
  u.o0 = __byte_perm(v.x, 0u, 0x4341);   // bytes 1,3
  u.e1 = __byte_perm(v.y, 0u, 0x4240);
  u.o1 = __byte_perm(v.y, 0u, 0x4341);
  return u;
}
__device__ __forceinline__ uint2 pack_u8x8(const U16x8& u) {
  return make_uint2(__byte_perm(u.e0, u.o0, 0x6240), __byte_perm(u.e1, u.o1, 0x6240));
}
__device__ __forceinline__ void max_u16x8(U16x8& a, const U16x8& b) {
  a.e0 = __vmaxu2(a.e0, b.e0);
  a.o0 = __vmaxu2(a.o0, b.o0);
  a.e1 = __vmaxu2(a.e1, b.e1);
  a.o1 = __vmaxu2(a.o1, b.o1);
}
__device__ __forceinline__ void max3_u16x8(U16x8& a, const U16x8& b, const U16x8& c) {
  a.e0 = __vimax3_u16x2(a.e0, b.e0, c.e0);
  a.o0 = __vimax3_u16x2(a.o0, b.o0, c.o0);
  a.e1 = __vimax3_u16x2(a.e1, b.e1, c.e1);
  a.o1 = __vimax3_u16x2(a.o1, b.o1, c.o1);
}
// max of the 8 values -> [0,255]
__device__ __forceinline__ uint32_t hmax_u16x8(const U16x8& u) {
  uint32_t q = __vimax3_u16x2(u.e0, u.o0, u.e1);
  q = __vmaxu2(q, u.o1);
  return __vmaxu2(q, q >> 16) & 0xffffu;
}

template <typename OutT, int NR>
__device__ __forceinline__ void k1u8_row_warp(const K1U8Params& p, const unsigned char* slabs,
                                              OutT* stg0, uint64_t* full, uint64_t* empty,
                                              uint64_t* done, uint64_t* sfree, uint32_t* norm_acc,
                                              int warp, int lane) {
  constexpr int kStgBytes = k1_staging_bytes<OutT>();
  const int off_yz = (p.mask & 1u) ? kFxz : 0;
  const int off_xy = off_yz + ((p.mask & 2u) ? kFyz : 0);
  const int j0 = warp * 4;
  const bool act = lane < kU8RowLanes;
  U16x8 yz[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) yz[r] = U16x8{0u, 0u, 0u, 0u};
  uint32_t sumsq = 0;
  uint32_t it = 0, t = 0;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++t) {
    const int buf = t & 1;
    OutT* stg = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(stg0) + buf * kStgBytes);
    mbar_wait(&sfree[buf], ((t >> 1) & 1) ^ 1);
    // two slabs per step (22 is even and the ring depth is even, so a pair never straddles a scan
    // or the ring end): the running max takes both with one three-input maximum per register
    for (int i = 0; i < kSX; i += 2, it += 2) {
      const int stage = it % kU8Stages;
      const uint32_t ph = (it / kU8Stages) & 1;
      mbar_wait(&full[stage], ph);
      mbar_wait(&full[stage + 1], ph);
      const uint2* rows = reinterpret_cast<const uint2*>(slabs + stage * kSlabBytesU8 + j0 * kSZ);
      constexpr int kNext = kSlabBytesU8 / 8;      // uint2 per slab
      uint2 va[NR], vb[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        va[r] = act ? rows[r * kU8RowLanes + lane] : make_uint2(0u, 0u);
        vb[r] = act ? rows[kNext + r * kU8RowLanes + lane] : make_uint2(0u, 0u);
      }
      uint32_t ma[NR], mb[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const U16x8 ua = unpack_u8x8(va[r]);
        const U16x8 ub = unpack_u8x8(vb[r]);
        max3_u16x8(yz[r], ua, ub);
        ma[r] = hmax_u16x8(ua);
        mb[r] = hmax_u16x8(ub);
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        ma[r] = __reduce_max_sync(0xffffffffu, ma[r]);
        mb[r] = __reduce_max_sync(0xffffffffu, mb[r]);
      }
      __syncwarp();
      if (lane == 0) {                               // both slabs have been consumed
        mbar_arrive(&empty[stage]);
        mbar_arrive(&empty[stage + 1]);
      }
      // lanes 0..NR-1 store the xy values of slab i, lanes NR..2NR-1 those of slab i+1
      uint32_t mine = 0;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        mine = (lane == r) ? ma[r] : mine;
        mine = (lane == NR + r) ? mb[r] : mine;
      }
      if (lane < 2 * NR && (p.mask & 4u)) {
        const int second = lane >= NR ? 1 : 0;
        EmitB<OutT>::put1(stg, off_xy + (i + second) * kSY + j0 + lane - second * NR, mine, p, sumsq);
      }
    }
    // end of scan: the running max over i is the yz projection
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (act && (p.mask & 2u))
        EmitB<OutT>::put8(stg, off_yz + (j0 + r) * kSZ + 8 * lane, pack_u8x8(yz[r]), p, sumsq);
      yz[r] = U16x8{0u, 0u, 0u, 0u};
    }
    if (sizeof(OutT) == 1) {
      const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
      if (lane == 0 && tot) atomicAdd(&norm_acc[buf], tot);
      sumsq = 0;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[buf]);
  }
}

// u8 rows: 108 KB of shared memory and 48 registers per thread, so two CTAs share an SM — the
// kernel is issue-bound (no byte-wise max instruction), and the second CTA fills the issue slots
// the first leaves idle while it waits on barriers.  f32 rows (167 KB) run one CTA per SM.
template <typename OutT>
__host__ __device__ constexpr int k1u8_ctas_per_sm() { return sizeof(OutT) == 1 ? 2 : 1; }

template <typename OutT>
__global__ void __launch_bounds__(kK1Threads, k1u8_ctas_per_sm<OutT>()) k1_project_max_u8in(const K1U8Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* slabs = smem;
  OutT* stg0 = reinterpret_cast<OutT*>(smem + kU8Stages * kSlabBytesU8);
  constexpr int kStgBytes = k1_staging_bytes<OutT>();
  unsigned char* tail = smem + kU8Stages * kSlabBytesU8 + 2 * kStgBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);   // [kU8Stages]
  uint64_t* empty = full + kU8Stages;                   // [kU8Stages]
  uint64_t* done = empty + kU8Stages;                   // [2] scan finished in staging buf
  uint64_t* sfree = done + 2;                           // [2] staging buf flushed
  uint32_t* norm_acc = reinterpret_cast<uint32_t*>(sfree + 2);  // [2]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kU8Stages; ++s) {
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
  for (int i = threadIdx.x; i < 2 * kStgBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(stg0)[i] = 0u;
  __syncthreads();

  uint32_t it = 0;  // running slab counter (identical in every role)
  uint32_t t = 0;   // running scan counter of this CTA

  if (warp < kRowWarps) {
    if (warp == kRowWarps - 1)
      k1u8_row_warp<OutT, kSY - 4 * (kRowWarps - 1)>(p, slabs, stg0, full, empty, done, sfree, norm_acc, warp, lane);
    else
      k1u8_row_warp<OutT, 4>(p, slabs, stg0, full, empty, done, sfree, norm_acc, warp, lane);
  } else if (warp < kRowWarps + kColWarps) {
    // ------------------------------------------------------------------ column warps
    const int c = warp - kRowWarps;
    const bool act = lane < kU8RowLanes;
    uint32_t sumsq = 0;
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++t) {
      const int buf = t & 1;
      OutT* stg = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(stg0) + buf * kStgBytes);
      mbar_wait(&sfree[buf], ((t >> 1) & 1) ^ 1);
      for (int i = 0; i < kSX; ++i, ++it) {
        if ((it % kColWarps) != static_cast<uint32_t>(c)) continue;   // ownership follows the ring slot
        const int stage = it % kU8Stages;
        mbar_wait(&full[stage], (it / kU8Stages) & 1);
        if ((p.mask & 1u) && act) {
          const uint2* slab = reinterpret_cast<const uint2*>(slabs + stage * kSlabBytesU8);
          // 31 rows: one alone, then 15 pairs folded by three-input maxima
          U16x8 m = unpack_u8x8(slab[lane]);
#pragma unroll
          for (int j = 1; j < kSY; j += 2)
            max3_u16x8(m, unpack_u8x8(slab[j * kU8RowLanes + lane]),
                       unpack_u8x8(slab[(j + 1) * kU8RowLanes + lane]));
          EmitB<OutT>::put8(stg, i * kSZ + 8 * lane, pack_u8x8(m), p, sumsq);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      if (sizeof(OutT) == 1) {
        const uint32_t tot = __reduce_add_sync(0xffffffffu, sumsq);
        if (lane == 0 && tot) atomicAdd(&norm_acc[buf], tot);
        sumsq = 0;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[buf]);
    }
  } else if (warp == kRowWarps + kColWarps) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        const uint8_t* cube = p.cubes + b * kCubeElems;
        for (int i = 0; i < kSX; ++i, ++it) {
          const int stage = it % kU8Stages;
          mbar_wait(&empty[stage], ((it / kU8Stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[stage], kSlabBytesU8);
          bulk_g2s(slabs + stage * kSlabBytesU8, cube + i * kSlabBytesU8, kSlabBytesU8, &full[stage], pol);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ flusher
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++t) {
      const int buf = t & 1;
      OutT* stg = reinterpret_cast<OutT*>(reinterpret_cast<unsigned char*>(stg0) + buf * kStgBytes);
      mbar_wait(&done[buf], (t >> 1) & 1);
      if (sizeof(OutT) == 1) {
        if (lane == 0) {
          bulk_s2g(reinterpret_cast<uint8_t*>(p.feats) + b * p.stride, stg, p.stride);
          bulk_commit();
          if (p.norms) p.norms[b] = static_cast<int32_t>(norm_acc[buf]);
          norm_acc[buf] = 0;
          bulk_wait_read<0>();
          mbar_arrive(&sfree[buf]);
          if (p.tile_done) {
            bulk_wait_all<0>();          // the feature row has landed in global memory
            __threadfence();
            atomicAdd(&p.tile_done[b >> 7], 1u);
          }
        }
      } else {
        float* out = reinterpret_cast<float*>(p.feats) + b * static_cast<int64_t>(p.stride);
        const float* src = reinterpret_cast<const float*>(stg);
        for (int idx = 2 * lane; idx < p.F; idx += 64)
          *reinterpret_cast<float2*>(out + idx) = *reinterpret_cast<const float2*>(src + idx);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfree[buf]);
      }
    }
    if (sizeof(OutT) == 1 && lane == 0) bulk_wait_all<0>();
  }
}

}  // namespace rml
