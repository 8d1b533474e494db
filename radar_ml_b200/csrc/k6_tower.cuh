// K6 — the conv towers of dnn.py / sgan.py fused on chip (SURVEY.md §8a A12-A14, §7 hard parts).
//
//   k6_tower<C1, FUSE2>   one persistent CTA per SM, bound to one branch (xz | yz | xy):
//     PIL resize (bit-exact, k3_net.cuh arithmetic)  ->  R x R image in shared memory
//     layer 1  Conv2D(1 -> C1, 3x3, s2, 'same') + bias + ReLU/LeakyReLU          dnn.py:48 / sgan.py:136-141
//              as a tcgen05 GEMM: A = im2col rows [128 pixels][K = 32] built in shared memory,
//              the fp32 image split hi + lo into bf16 and the fp32 kernel split the same way so
//              that hi.whi + lo.whi + hi.wlo keeps ~16 bits more than a plain bf16 product (the
//              reference layer is fp32); accumulators in TMEM, four 128-pixel tiles in flight
//     FUSE2    layer 2  Conv2D(64 -> C2, 3x3, s2, 'same') + bias + act            dnn.py:50
//              straight from shared memory: the bf16 layer-1 rows of a strip of output rows are
//              written by the layer-1 epilogue as four parity planes [row parity][col parity] of
//              128-byte pixel rows in the 128B-swizzled K-major layout tcgen05.mma reads, so each
//              of the nine taps is ONE descriptor (plane base + (kh>>1) rows + (kw>>1) pixels) —
//              the layer-1 activation (614 KB per scan in bf16) never touches HBM or L2
//     !FUSE2   the layer-1 rows go to global memory as NHWC bf16 (sgan: 128 channels do not fit
//              the on-chip strip next to the 147 KB layer-2 kernel; layers 2-3 stay k4_conv_igemm)
//
// Measured facts this relies on (tools/umma_probe.cu on a B200): the 128B swizzle is a function of
// the absolute shared-memory address, so a descriptor may start at any 128-byte aligned row of an
// array stored with that swizzle (base_offset field 0); the no-swizzle K-major descriptor takes
// LBO = K-direction core-matrix stride, SBO = M-direction 8-row group stride.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "k3_net.cuh"
#include "ptx.cuh"

namespace rml {

constexpr int kT6Threads = 512;          // 4 groups of 4 warps: a group owns one 128-pixel tile
constexpr int kT6K1 = 32;                // layer-1 GEMM K: 9 hi | 9 lo | 9 hi | 5 zero
constexpr int kT6A1Bytes = 128 * kT6K1 * 2;   // 8 KB im2col tile, no-swizzle K-major
constexpr int kT6TH = 5;                 // layer-2 output rows per strip (FUSE2)
constexpr int kT6Pitch = 24;             // plane slots per row: 20 pixels + the zero pad slot, 8-aligned

struct TowerParams {
  ResizeParams rz;               // rz.images unused; rz.feats = [B][F] scaled projections
  const float* images;           // nullable: [B][3][R][R] already resized inputs (Keras model.predict
                                 // convention, dnn.py:371-381) — the resize is skipped, same towers
  int64_t B;
  int ctas[3];                   // persistent CTAs per branch (sum = gridDim.x)
  const __nv_bfloat16* w1;       // [3][C1][32] layer-1 kernel rows: whi(9) | whi(9) | wlo(9) | bias hi | bias lo | 0(3)
                                 // (the bias rides in the GEMM against two 1.0 columns of the im2col row)
  const __nv_bfloat16* w2;       // FUSE2: [3][C2][9*64] bf16, K = tap*64 + ci (k4_conv_igemm layout)
  const float* b2;               // FUSE2: [3][C2]
  float alpha;                   // LeakyReLU slope (ACT == 2)
  __nv_bfloat16* out;            // FUSE2: [B][3][H2][H2][C2]; else [B*3][H1][H1][C1]
};

__host__ __device__ constexpr int t6_img_pitch(int R) { return R + 4; }
// shared-memory plan (bytes); every region 1024-aligned where tcgen05 reads it
template <int C1, bool FUSE2>
struct T6Smem {
  static constexpr int R = FUSE2 ? 80 : 128;
  static constexpr int C2 = 32;
  static constexpr int src = 0;                                          // projection [H][W] fp32 (<= 31 x 176)
  static constexpr int img = src + 31 * 176 * 4;                         // [(R+1)][R+4] fp32, zero row / col R
  static constexpr int a1 = (img + (R + 1) * t6_img_pitch(R) * 4 + 1023) & ~1023;   // 4 im2col tiles
  static constexpr int w1 = a1 + 4 * kT6A1Bytes;                         // [C1][32] bf16 no-swizzle K-major
  static constexpr int w2 = (w1 + C1 * kT6K1 * 2 + 1023) & ~1023;        // FUSE2: 9 taps x [C2][64] bf16, SW128
  static constexpr int strip = w2 + (FUSE2 ? 9 * C2 * 128 : 0);          // FUSE2: 4 parity planes + tail
  static constexpr int strip_slots = 2 * (kT6TH + 1) * kT6Pitch + 2 * kT6TH * kT6Pitch + 32;
  // resize scratch in float64 (the projection and the horizontal pass converted ONCE instead of once
  // per tap): it is dead before the first strip is written, so FUSE2 lays it over the strip planes
  static constexpr int srcd_bytes = 31 * 176 * 8, tmpd_bytes = 31 * R * 8;
  static constexpr int strip_bytes = FUSE2 ? strip_slots * 128 : 0;
  static constexpr int scratch = strip;                                   // srcd | tmpd
  static constexpr int scratch_bytes = (srcd_bytes + tmpd_bytes > strip_bytes) ? srcd_bytes + tmpd_bytes : strip_bytes;
  static constexpr int tabs = strip + scratch_bytes;                      // kh [R][12] | kv [R][8] f64, bh | bv [R] int2
  static constexpr int tabs_bytes = R * (12 + 8) * 8 + 2 * R * 8;
  static constexpr int bias = tabs + tabs_bytes;                          // b2 [C2]
  static constexpr int bars = bias + C2 * 4;
  static constexpr int total = bars + 128 + 1024;                        // + alignment slack
  static_assert(total <= 232448, "shared memory plan exceeds 227 KB");
};

__device__ __forceinline__ uint64_t t6_desc_noswz(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 16;     // LBO: next core matrix along K
  d |= static_cast<uint64_t>(512 >> 4) << 32;     // SBO: next 8-row group along M / N
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bar_group(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Pillow's two resample passes (Resample.c ImagingResampleHorizontal/Vertical_32bpc) with the same
// arithmetic as k3_resize_pil — ss = 0.0; ss += (double)in[i] * k[i] in tap order; store (float)ss —
// hence bit-identical, but every input is converted to float64 once: the projection when it is
// copied to `srcd`, the horizontal result when it is stored ((double)(float)ss).  Coefficient
// tables come from shared memory; a thread owns one output column and two rows per step (two
// independent DFMA chains).
template <int KT>
__device__ __forceinline__ void t6_pass_h(const double* srcd, double* tmpd, int H, int W, int R,
                                          const double* s_kh, const int2* s_bh) {
  const int groups = kT6Threads / R;
  const int xx = threadIdx.x % R, g = threadIdx.x / R;
  if (g >= groups) return;
  const int2 bd = s_bh[xx];
  double k[KT];
#pragma unroll
  for (int x = 0; x < KT; ++x) k[x] = x < bd.y ? s_kh[xx * 12 + x] : 0.0;
  for (int y = g; y < H; y += 2 * groups) {
    const int y2 = y + groups;
    const bool two = y2 < H;
    const double* in0 = srcd + y * W + bd.x;
    const double* in1 = srcd + (two ? y2 : y) * W + bd.x;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int x = 0; x < KT; ++x)
      if (x < bd.y) {
        s0 += in0[x] * k[x];
        s1 += in1[x] * k[x];
      }
    tmpd[y * R + xx] = static_cast<double>(static_cast<float>(s0));
    if (two) tmpd[y2 * R + xx] = static_cast<double>(static_cast<float>(s1));
  }
}
template <int KT>
__device__ __forceinline__ void t6_pass_v(const double* tmpd, float* img, int P, int R,
                                          const double* s_kv, const int2* s_bv) {
  const int groups = kT6Threads / R;
  const int xx = threadIdx.x % R, g = threadIdx.x / R;
  if (g >= groups) return;
  for (int yy = g; yy < R; yy += 2 * groups) {
    const int yy2 = yy + groups;
    const bool two = yy2 < R;
    const int2 b0 = s_bv[yy], b1 = s_bv[two ? yy2 : yy];
    const double* in0 = tmpd + b0.x * R + xx;
    const double* in1 = tmpd + b1.x * R + xx;
    const double* k0 = s_kv + yy * 8;
    const double* k1 = s_kv + (two ? yy2 : yy) * 8;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int y = 0; y < KT; ++y) {
      if (y < b0.y) s0 += in0[y * R] * k0[y];
      if (y < b1.y) s1 += in1[y * R] * k1[y];
    }
    img[yy * P + xx] = static_cast<float>(s0);
    if (two) img[yy2 * P + xx] = static_cast<float>(s1);
  }
}

// relu / leaky-relu on two fp32 values, packed to bf16x2
template <int ACT>
__device__ __forceinline__ uint32_t act_pack(float a, float b, float alpha) {
  uint32_t r;
  if (ACT == 1) {
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  } else {
    if (ACT == 2) { a = fmaxf(a, alpha * a); b = fmaxf(b, alpha * b); }   // alpha < 1
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  }
  return r;
}

template <int C1, bool FUSE2, int ACT>
__global__ void __launch_bounds__(kT6Threads, 1) k6_tower(const TowerParams p) {
  using L = T6Smem<C1, FUSE2>;
  constexpr int R = L::R;
  constexpr int H1 = R / 2;               // layer-1 output size
  constexpr int C2 = L::C2;
  constexpr int P = t6_img_pitch(R);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* src = reinterpret_cast<float*>(smem + L::src);
  float* img = reinterpret_cast<float*>(smem + L::img);
  double* srcd = reinterpret_cast<double*>(smem + L::scratch);
  double* tmpd = reinterpret_cast<double*>(smem + L::scratch + L::srcd_bytes);
  double* s_kh = reinterpret_cast<double*>(smem + L::tabs);            // [R][12]
  double* s_kv = s_kh + R * 12;                                         // [R][8]
  int2* s_bh = reinterpret_cast<int2*>(s_kv + R * 8);
  int2* s_bv = s_bh + R;
  unsigned char* a1 = smem + L::a1;
  unsigned char* w1s = smem + L::w1;
  unsigned char* w2s = smem + L::w2;
  unsigned char* strip = smem + L::strip;
  float* s_b2 = reinterpret_cast<float*>(smem + L::bias);
  uint64_t* mbar1 = reinterpret_cast<uint64_t*>(smem + L::bars);   // [4] layer-1 tile of group g done
  uint64_t* mbar2 = mbar1 + 4;                                      // layer-2 strip done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar2 + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2;               // tile group 0..3
  const int q = warp & 3;                  // TMEM lane quarter of this warp
  const int m = q * 32 + lane;             // row of the group's tile

  // which branch, which scans
  int br = 0, first = blockIdx.x;
  if (first >= p.ctas[0]) { first -= p.ctas[0]; br = 1; }
  if (br == 1 && first >= p.ctas[1]) { first -= p.ctas[1]; br = 2; }
  const int n_mine = p.ctas[br];
  const int H = p.rz.ph[br], W = p.rz.pw[br];
  const double* kh = p.rz.kh[br];
  const double* kv = p.rz.kv[br];
  const int2* bh = p.rz.bh[br];
  const int2* bv = p.rz.bv[br];
  const int ksh = p.rz.ksh[br], ksv = p.rz.ksv[br];

  // ---- one-time setup: zero image border, weights into operand layouts, barriers, TMEM
  for (int e = tid; e < (R + 1) * P; e += kT6Threads) img[e] = 0.f;
  for (int e = tid; e < C1 * 4; e += kT6Threads) {          // 16-byte chunk (n, kc) of the layer-1 kernel
    const int n = e >> 2, kc = e & 3;
    *reinterpret_cast<uint4*>(w1s + (n >> 3) * 512 + kc * 128 + (n & 7) * 16) =
        *reinterpret_cast<const uint4*>(p.w1 + (static_cast<size_t>(br) * C1 + n) * kT6K1 + kc * 8);
  }
  for (int e = tid; e < R * 12; e += kT6Threads) { const int xx = e / 12, x = e - xx * 12; s_kh[e] = x < ksh ? kh[xx * ksh + x] : 0.0; }
  for (int e = tid; e < R * 8; e += kT6Threads) { const int yy = e >> 3, y = e & 7; s_kv[e] = y < ksv ? kv[yy * ksv + y] : 0.0; }
  for (int e = tid; e < R; e += kT6Threads) { s_bh[e] = bh[e]; s_bv[e] = bv[e]; }
  if (FUSE2) {
    for (int e = tid; e < 9 * C2 * 8; e += kT6Threads) {    // chunk c of row n of tap t
      const int c = e & 7, n = (e >> 3) % C2, t = e / (8 * C2);
      *reinterpret_cast<uint4*>(w2s + t * (C2 * 128) + n * 128 + ((c ^ (n & 7)) << 4)) =
          *reinterpret_cast<const uint4*>(p.w2 + (static_cast<size_t>(br) * C2 + n) * 576 + t * 64 + c * 8);
    }
    for (int e = tid; e < C2; e += kT6Threads) s_b2[e] = p.b2[br * C2 + e];
    for (int e = tid; e < L::strip_slots * 32; e += kT6Threads) reinterpret_cast<uint32_t*>(strip)[e] = 0u;
  }
  if (tid == 0) {
    for (int g = 0; g < 4; ++g) mbar_init(&mbar1[g], 1);
    mbar_init(mbar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem1 = tmem_base + grp * C1;                // layer-1 accumulator of this group (C1 <= 128)
  const uint32_t tmem2 = tmem_base + (FUSE2 ? 4 * C1 : 0);    // layer-2 accumulator (FUSE2: C1 = 64 -> column 256)
  const uint32_t idesc1 = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, C1);
  const uint32_t idesc2 = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, C2);
  uint32_t ph1 = 0, ph2 = 0;               // mbarrier phases (per thread copies stay in step)

  // prefetch the first projection
  auto prefetch = [&](int64_t b) {
    const float* g = p.rz.feats + b * p.rz.F + p.rz.poff[br];       // 8-byte aligned rows
    for (int e = tid; e < (H * W) / 2; e += kT6Threads) cp_async8(src + 2 * e, g + 2 * e);
  };
  if (first < p.B && !p.images) prefetch(first);

  // plane offsets (FUSE2): [py][px] planes of (TH+1 | TH) rows x kT6Pitch slots x 128 B
  constexpr int kPlane0 = (kT6TH + 1) * kT6Pitch * 128, kPlane1 = kT6TH * kT6Pitch * 128;

  for (int64_t b = first; b < p.B; b += n_mine) {
    if (p.images) {
      const float* g = p.images + (static_cast<int64_t>(b) * 3 + br) * R * R;
      for (int e = tid; e < R * R; e += kT6Threads) img[(e / R) * P + (e % R)] = g[e];
      __syncthreads();
    } else {
      cp_async_wait_all();
      __syncthreads();
      for (int e = tid; e < H * W; e += kT6Threads) srcd[e] = static_cast<double>(src[e]);
      __syncthreads();
      if (b + n_mine < p.B) prefetch(b + n_mine);              // src is free from here on
      if (ksh <= 5) t6_pass_h<5>(srcd, tmpd, H, W, R, s_kh, s_bh);
      else if (ksh <= 8) t6_pass_h<8>(srcd, tmpd, H, W, R, s_kh, s_bh);
      else t6_pass_h<12>(srcd, tmpd, H, W, R, s_kh, s_bh);
      __syncthreads();
      if (ksv <= 5) t6_pass_v<5>(tmpd, img, P, R, s_kv, s_bv);  // TF 'same' pads after: the image starts at (0,0)
      else t6_pass_v<8>(tmpd, img, P, R, s_kv, s_bv);
      __syncthreads();
      if (FUSE2) {
        // the scratch lay over the strip planes: the zero pad slots (pixel 20 of every even-column
        // plane row) must read as zero again before layer 2 touches them
        constexpr int kPl0 = (kT6TH + 1) * kT6Pitch * 128, kPl1 = kT6TH * kT6Pitch * 128;
        for (int e = tid; e < (2 * kT6TH + 1) * 4 * 8; e += kT6Threads) {
          const int c = e & 7, sl = (e >> 3) & 3, rr = e >> 5;      // chunk, slot 20..23, plane row
          const int py = rr > kT6TH, a = py ? rr - (kT6TH + 1) : rr;
          unsigned char* row = strip + (py ? 2 * kPl0 : 0) + (a * kT6Pitch + kT6Pitch - 4 + sl) * 128;
          *reinterpret_cast<uint4*>(row + c * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        (void)kPl1;
      }
    }

    // ------------------------------------------------------------------ layer 1 (+ layer 2) over tiles
    constexpr int kRowsPerStrip = FUSE2 ? 2 * kT6TH + 1 : 8;             // layer-1 rows per round of 4 tiles
    constexpr int kPixPerStrip = kRowsPerStrip * H1;                     // 440 (dnn) / 512 (sgan)
    constexpr int kStrips = FUSE2 ? (H1 / 2) / kT6TH : H1 / 8;           // 4 / 8
    static_assert(kPixPerStrip <= 512, "four 128-pixel tiles per round");
    for (int s = 0; s < kStrips; ++s) {
      const int y1_0 = FUSE2 ? 2 * kT6TH * s : 8 * s;       // first layer-1 row of the round
      const int pidx = grp * 128 + m;                        // pixel of this thread inside the round
      const int r = pidx / H1, x1 = pidx - r * H1;
      const int y1 = y1_0 + r;
      const bool live = pidx < kPixPerStrip && y1 < H1;      // y1 == H1 is layer 2's zero pad row
      // -- im2col row of this pixel: 9 taps, hi / lo split
      {
        float v[9];
        if (live) {
          const float* ip = img + (2 * y1) * P + 2 * x1;
#pragma unroll
          for (int t3 = 0; t3 < 3; ++t3) {
            v[t3 * 3 + 0] = ip[t3 * P + 0];
            v[t3 * 3 + 1] = ip[t3 * P + 1];
            v[t3 * 3 + 2] = ip[t3 * P + 2];
          }
        } else {
#pragma unroll
          for (int t = 0; t < 9; ++t) v[t] = 0.f;
        }
        float l[9];
        unsigned short hb[9], lb[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[t]);
          l[t] = v[t] - __bfloat162float(h);
          hb[t] = __bfloat16_as_ushort(h);
          lb[t] = __bfloat16_as_ushort(__float2bfloat16_rn(l[t]));
        }
        // K order: hi0..8 | lo0..8 | hi0..8 | 1 | 1 | 0 x 3   (the 1.0 columns meet bias hi / lo in the
        // kernel rows; a dead pixel keeps an all-zero row, so its output is exactly 0 = the pad value)
        unsigned short kk[32];
#pragma unroll
        for (int t = 0; t < 9; ++t) { kk[t] = hb[t]; kk[9 + t] = lb[t]; kk[18 + t] = hb[t]; }
        kk[27] = kk[28] = live ? 0x3F80 : 0;      // bf16(1.0)
#pragma unroll
        for (int t = 29; t < 32; ++t) kk[t] = 0;
        unsigned char* row = a1 + grp * kT6A1Bytes + (m >> 3) * 512 + (m & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint4 w;
          w.x = kk[kc * 8 + 0] | (static_cast<uint32_t>(kk[kc * 8 + 1]) << 16);
          w.y = kk[kc * 8 + 2] | (static_cast<uint32_t>(kk[kc * 8 + 3]) << 16);
          w.z = kk[kc * 8 + 4] | (static_cast<uint32_t>(kk[kc * 8 + 5]) << 16);
          w.w = kk[kc * 8 + 6] | (static_cast<uint32_t>(kk[kc * 8 + 7]) << 16);
          *reinterpret_cast<uint4*>(row + kc * 128) = w;
        }
      }
      fence_proxy_async_smem();
      bar_group(1 + grp, 128);
      // -- layer-1 GEMM of this group's tile: 128 pixels x C1 channels, K = 32 (two UMMAs)
      if (q == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(a1 + grp * kT6A1Bytes);
          const uint32_t b_addr = smem_u32(w1s);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f16(tmem1, t6_desc_noswz(a_addr + k * 256), t6_desc_noswz(b_addr + k * 256), idesc1, k != 0);
          umma_commit(&mbar1[grp]);
        }
        __syncwarp();
      }
      mbar_wait(&mbar1[grp], ph1);
      ph1 ^= 1;
      tc_fence_after();
      if (FUSE2 && s > 0) {
        // the strip buffer is still being read by the layer-2 MMAs of the previous strip
        mbar_wait(mbar2, ph2 ^ 1);
      }
      // -- layer-1 epilogue: bias + activation -> bf16 -> strip planes (FUSE2) or global NHWC
      {
        const uint32_t taddr = tmem1 + (static_cast<uint32_t>(q * 32) << 16);
        unsigned char* dst_row = nullptr;
        uint32_t sw = 0;
        __nv_bfloat16* gout = nullptr;
        if (FUSE2) {
          const int py = r & 1, a = r >> 1, px = x1 & 1, j = x1 >> 1;
          const int slot = a * kT6Pitch + j;
          dst_row = strip + (py ? 2 * kPlane0 + px * kPlane1 : px * kPlane0) + slot * 128;
          sw = slot & 7;
        } else {
          gout = p.out + ((static_cast<int64_t>(b) * 3 + br) * H1 * H1 + static_cast<int64_t>(y1) * H1 + x1) * C1;
        }
        const bool store = pidx < kPixPerStrip && (FUSE2 || y1 < H1);
#pragma unroll
        for (int c0 = 0; c0 < C1; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2)
            pk[e >> 1] = act_pack<ACT>(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), p.alpha);
          if (store) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 w = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
              if (FUSE2) *reinterpret_cast<uint4*>(dst_row + ((((c0 >> 3) + c) ^ sw) << 4)) = w;
              else *reinterpret_cast<uint4*>(gout + c0 + 8 * c) = w;
            }
          }
        }
      }
      tc_fence_before();
      if (!FUSE2) {
        bar_group(1 + grp, 128);          // the tile's TMEM and im2col buffer are free for the next round
        continue;
      }
      fence_proxy_async_smem();
      __syncthreads();                    // the whole strip is in shared memory
      // -- layer 2: 9 taps x 4 UMMAs (K = 64 channels) straight from the parity planes
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sbase = smem_u32(strip);
          const uint32_t wbase = smem_u32(w2s);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int kh3 = t / 3, kw3 = t - 3 * kh3;
            const uint32_t a_addr = sbase + ((kh3 & 1) ? 2 * kPlane0 + (kw3 & 1) * kPlane1 : (kw3 & 1) * kPlane0) +
                                    ((kh3 >> 1) * kT6Pitch + (kw3 >> 1)) * 128;
            const uint64_t da = umma_desc_k_sw128(a_addr);
            const uint64_t db = umma_desc_k_sw128(wbase + t * (C2 * 128));
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16(tmem2, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc2, (t | ks) != 0);
          }
          umma_commit(mbar2);
        }
        __syncwarp();
      }
      ph2 ^= 1;                            // every thread tracks the phase; group 0 waits for it now
      if (grp == 0) {
        mbar_wait(mbar2, ph2 ^ 1);
        tc_fence_after();
        const int oyl = m / kT6Pitch, ox = m - oyl * kT6Pitch;
        const int oy = kT6TH * s + oyl;
        const uint32_t taddr = tmem2 + (static_cast<uint32_t>(q * 32) << 16);
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (oyl < kT6TH && ox < H1 / 2) {
          __nv_bfloat16* o = p.out + (((static_cast<int64_t>(b) * 3 + br) * (H1 / 2) + oy) * (H1 / 2) + ox) * C2;
#pragma unroll
          for (int c = 0; c < C2 / 8; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              pk[e] = act_pack<ACT>(__uint_as_float(v[8 * c + 2 * e]) + s_b2[8 * c + 2 * e],
                                    __uint_as_float(v[8 * c + 2 * e + 1]) + s_b2[8 * c + 2 * e + 1], p.alpha);
            reinterpret_cast<uint4*>(o)[c] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        tc_fence_before();
      }
    }
    if (FUSE2) {
      // the last strip's MMAs must retire before the next image's epilogue overwrites the planes;
      // group 0 has already waited, the others catch up here
      if (grp != 0) mbar_wait(mbar2, ph2 ^ 1);
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace rml
