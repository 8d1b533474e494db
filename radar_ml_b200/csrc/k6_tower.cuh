// K6 — the conv towers of dnn.py / sgan.py fused on chip (SURVEY.md §8a A12-A14, §7 hard parts).
//
//   k6_tower<C1, FUSE2, ACT>   one persistent CTA per SM, bound to one branch (xz | yz | xy), 24 warps
//   in two roles that run concurrently:
//     R role (8 warps)   PIL resize, bit-exact (k3_net.cuh arithmetic): projection (cp.async prefetch)
//                        -> horizontal pass -> vertical pass, published to the T role block of rows
//                        by block of rows through full/empty mbarriers on the R x R image buffer
//                        (FUSE2: two image buffers, so the R role runs up to an image ahead)
//     T role (16 warps)  layer 1  Conv2D(1 -> C1, 3x3, s2, 'same') + bias + ReLU/LeakyReLU  dnn.py:48 / sgan.py:136-141
//                        as a tcgen05 GEMM: A = im2col rows [128 pixels][K = 32] — the fp32 image and
//                        kernel split hi + lo into fp16 pairs (22 bits, subnormals keep the absolute
//                        error at 2^-25) so that hi.whi + lo.whi + hi.wlo reproduces the fp32 layer to
//                        ~2^-21; the bias rides in two 1.0 columns; accumulators in TMEM, four
//                        128-pixel tiles in flight (one per 4-warp group)
//       FUSE2 (dnn)      the im2col rows are written straight into TENSOR MEMORY (tcgen05.st, lane =
//                        pixel) and the layer-1 UMMAs take A from there (no shared-memory store or read);
//                        layer 2  Conv2D(64 -> C2, 3x3, s2, 'same') + bias + act        dnn.py:50
//                        straight from shared memory: the bf16 layer-1 rows of a strip of output rows
//                        are written by the layer-1 epilogue as four parity planes [row parity][col
//                        parity] of 128-byte pixel rows in the 128B-swizzled K-major layout tcgen05.mma
//                        reads, so each of the nine taps is ONE descriptor (plane base + (kh>>1) rows +
//                        (kw>>1) pixels) — the layer-1 activation (614 KB per scan in bf16) never
//                        touches HBM or L2.  Rounds are software-pipelined (see the loop) and the one
//                        warp without pixels issues every UMMA.
//       !FUSE2 (sgan)    im2col tiles in shared memory; the layer-1 rows go to global memory as NHWC bf16
//                        (128 channels do not fit the on-chip strip next to the 147 KB layer-2 kernel;
//                        layers 2-3 stay k4_conv_igemm).  That write (3.1 MB per scan) bounds sgan.
//
// What bounds it (profiles/r2_tower_variants.txt): with either role's arithmetic switched off the dnn
// towers take 1.20 ms (R only) / 1.52 ms (T only) per 8 192 scans, together 2.03 ms, and 0.81 ms with
// NO arithmetic at all — the two roles slow each other down on the shared load/store and issue paths
// (24 warps, ~1 instruction per 12-20 cycles per warp) and the per-round barrier / commit round trips
// (~2.4 k cycles) are in series with the work.  Shared-memory wavefronts were cut from 94 % to 76 % of
// the data pipe (A operand in TMEM, padded planes) without moving the time.
//
// Measured facts this relies on (tools/umma_probe.cu on a B200, profiles/r2_umma_probe.txt): the 128B
// swizzle is a function of the absolute shared-memory address, so a descriptor may start at any
// 128-byte aligned row of an array stored with that swizzle (base_offset field 0); the no-swizzle
// K-major descriptor takes LBO = K-direction core-matrix stride, SBO = M-direction 8-row group stride;
// a tcgen05.mma of N <= 64 costs ~32 + N/4 cycles whatever its floor.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "k3_net.cuh"
#include "ptx.cuh"

namespace rml {

constexpr int kT6TThreads = 512;         // T role: 4 groups of 4 warps, a group owns one 128-pixel tile
constexpr int kT6RThreads = 256;         // R role: 8 warps
constexpr int kT6Threads = kT6TThreads + kT6RThreads;
constexpr int kT6K1 = 32;                // layer-1 GEMM K: hi0..8, 1 | lo0..8, 0 | hi0..8, 1 | 0, 0
constexpr int kT6A1Bytes = 128 * kT6K1 * 2;   // 8 KB im2col tile, no-swizzle K-major
constexpr int kT6TH = 5;                 // layer-2 output rows per strip (FUSE2)
constexpr int kT6Pitch = 21;             // plane slots per row: 20 pixels + the zero pad slot

struct TowerParams {
  ResizeParams rz;               // rz.images unused; rz.feats = [B][F] scaled projections
  const float* images;           // nullable: [B][3][R][R] already resized inputs (Keras model.predict
                                 // convention, dnn.py:371-381) — the resize is skipped, same towers
  int64_t B;
  int ctas[3];                   // persistent CTAs per branch (sum = gridDim.x)
  const __half* w1;              // [3][C1][32] layer-1 kernel rows, fp16: whi(9), bias hi | whi(9), 0 | wlo(9), bias lo | 0, 0
  const __nv_bfloat16* w2;       // FUSE2: [3][C2][9*64] bf16, K = tap*64 + ci (k4_conv_igemm layout)
  const float* b2;               // FUSE2: [3][C2]
  float alpha;                   // LeakyReLU slope (ACT == 2)
  int dbg;                       // experiments only (RML_T6_DBG): 64 = strided-column horizontal pass, 32 = gather after
                                 // the layer-1 commit, 512 = barrier form of the mbarrier waits (same results);
                                 // 1 no layer-2 MMAs, 2 no resize, 4 no layer-1
                                 // epilogue, 8 no im2col, 16 no layer-1 MMAs — results are wrong with any bit set
  __nv_bfloat16* out;            // FUSE2: [B][3][H2][H2][C2]; else [B*3][H1][H1][C1]
};

__host__ __device__ constexpr int t6_img_pitch(int R) { return R + 5; }   // odd: row-per-lane stores are conflict-free
constexpr int kT6HP = 33;                // padded height of the transposed horizontal-pass buffer (odd, >= 31 + 2)
// shared-memory plan (bytes); every region 1024-aligned where tcgen05 reads it
template <int C1, bool FUSE2>
struct T6Smem {
  static constexpr int R = FUSE2 ? 80 : 128;
  static constexpr int C2 = 32;
  static constexpr int NB = FUSE2 ? 4 : 8;                               // image row blocks = T rounds per image
  static constexpr int NIMG = FUSE2 ? 2 : 1;                             // image buffers (FUSE2: the second one lives in the
                                                                         // im2col region — its A tiles are in tensor memory)
  static constexpr int src = 0;                                          // projection [H][W] fp32 (<= 31 x 176)
  static constexpr int tmpd = src + 31 * 176 * 4;                        // horizontal pass, transposed [R][33] float64
  static constexpr int img = tmpd + R * kT6HP * 8;                       // [(R+1)][R+5] fp32, zero row / col R
  static constexpr int a1 = (img + (R + 1) * t6_img_pitch(R) * 4 + 1023) & ~1023;   // 4 im2col tiles
  static_assert((R + 1) * t6_img_pitch(R) * 4 <= 4 * kT6A1Bytes || !FUSE2, "second image buffer");
  static constexpr int w1 = a1 + 4 * kT6A1Bytes;                         // [C1][32] fp16 no-swizzle K-major
  static constexpr int w2 = (w1 + C1 * kT6K1 * 2 + 1023) & ~1023;        // FUSE2: 9 taps x [C2][64] bf16, SW128
  static constexpr int strip = w2 + (FUSE2 ? 9 * C2 * 128 : 0);          // FUSE2: 4 parity planes + tail
  // plane sizes padded so that the even- and odd-column planes of a row parity start 4 slots apart
  // modulo 8: the 128-bit stores of a quarter-warp (4 even + 4 odd pixels) then hit 8 distinct bank
  // groups under the 128B swizzle (unpadded they collided two ways: 1 380 extra wavefronts per image)
  static constexpr int plane0_slots = ((kT6TH + 1) * kT6Pitch + 7) / 8 * 8 + 4, plane1_slots = (kT6TH * kT6Pitch + 7) / 8 * 8 - 4;
  static_assert(plane0_slots >= (kT6TH + 1) * kT6Pitch && plane1_slots >= kT6TH * kT6Pitch && plane0_slots % 8 == 4 && plane1_slots % 8 == 4, "plane padding");
  static constexpr int strip_slots = 2 * plane0_slots + 2 * plane1_slots + 24;     // tail: the last tile over-reads
  // !FUSE2: every T warp stages 32 pixels x 64 channels (4 KB, 128B-swizzled rows) for one TMA store; warps
  // 0-1 of a group use the group's im2col tile (free once its UMMAs have retired), warps 2-3 this region
  static constexpr int stg2 = (strip + (FUSE2 ? strip_slots * 128 : 0) + 1023) & ~1023;
  static constexpr int stg2_bytes = FUSE2 ? 0 : 4 * 2 * 4096;
  static constexpr int tabs = FUSE2 ? strip + strip_slots * 128 : stg2 + stg2_bytes;   // kh [R][12] | kv [R][8] f64, bh | bv [R] int2
  static constexpr int tabs_bytes = R * (12 + 8) * 8 + 2 * R * 8;
  static constexpr int bias = tabs + tabs_bytes;                          // b2 [C2]
  static constexpr int bars = bias + C2 * 4;
  static constexpr int total = bars + 256 + 1024;                        // + alignment slack
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
// long waits (another role's progress, a whole strip of MMAs): poll with a back-off so that the
// waiting warps leave the issue slots to the warps that are working
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(40);
}
__device__ __forceinline__ void bar_group(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Pillow's two resample passes (Resample.c ImagingResampleHorizontal/Vertical_32bpc) with the same
// arithmetic as k3_resize_pil — ss = 0.0; ss += (double)in[i] * k[i] in tap order; store (float)ss —
// hence bit-identical.  The horizontal result is kept as (double)(float)ss so the vertical pass
// converts nothing; coefficient tables come from shared memory.  `rt` = thread index in the R role.
//
// Horizontal: a thread owns one output column (its taps in registers) and walks down the rows, kT6Chains
// rows per step.  The resize role is LATENCY bound — 8 warps whose instructions wait behind the tile
// role's in the load/store queue (ncu: one instruction per ~20 cycles per warp) — so what counts is the
// number of independent load -> convert -> DFMA chains in flight per thread, not instructions or
// wavefronts.  (Measured alternatives, profiles/r2_tower_resize_variants.txt: lanes on every (R/16)-th
// column make every tap one conflict-free wavefront, -6 % wavefronts, but leave 3 of 8 warps idle: slower;
// a warp per column with lanes over the rows needs its coefficients as broadcast loads, 2 wavefronts per
// float64: no saving; the same from constant memory misses the constant cache on every dependent read.)
constexpr int kT6Chains = 2;   // 4 measured slower (1 605 against 1 489 us per 8 192 scans)
template <int KT>
__device__ __forceinline__ void t6_pass_h(int rt, bool strided, const float* src, double* tmpT, int H, int W, int R,
                                          const double* s_kh, const int2* s_bh) {
  int xx, y0, dy;
  if (strided) {                              // warp w: columns w + (R/16) j, half-warps on alternate rows (RML_T6_DBG=64)
    const int w = rt >> 5, lane = rt & 31, q = R >> 4;
    if (w >= q) return;
    xx = w + q * (lane & 15); y0 = lane >> 4; dy = 2;
  } else {                                    // neighbouring columns, 256 / R row groups (default)
    const int groups = kT6RThreads / R;
    xx = rt % R; y0 = rt / R; dy = groups;
    if (y0 >= groups) return;
  }
  const int2 bd = s_bh[xx];
  double k[KT];
#pragma unroll
  for (int x = 0; x < KT; ++x) k[x] = x < bd.y ? s_kh[x * R + xx] : 0.0;
  double* out = tmpT + xx * kT6HP;            // transposed: the vertical pass walks a column contiguously
  const float* in = src + bd.x;
  for (int y = y0; y < H; y += kT6Chains * dy) {      // rows y, y + dy, ...
    const float* ip[kT6Chains];
    double sum[kT6Chains];
#pragma unroll
    for (int c = 0; c < kT6Chains; ++c) {
      ip[c] = in + (y + c * dy < H ? y + c * dy : y) * W;
      sum[c] = 0.0;
    }
#pragma unroll
    for (int x = 0; x < KT; ++x)
      if (x < bd.y) {
#pragma unroll
        for (int c = 0; c < kT6Chains; ++c) sum[c] += static_cast<double>(ip[c][x]) * k[x];
      }
#pragma unroll
    for (int c = 0; c < kT6Chains; ++c)
      if (y + c * dy < H) out[y + c * dy] = static_cast<double>(static_cast<float>(sum[c]));
  }
}
// output rows [lo, hi] (at most 24) of the vertical pass: a thread owns ONE output row — its taps and
// bounds sit in registers — and walks over columns, kT6Chains at a time (independent DFMA chains);
// the lanes of a warp are consecutive rows, whose input windows overlap: the float64 loads of a
// step are near-broadcasts and the float stores hit distinct banks (odd image pitch)
template <int KT>
__device__ __forceinline__ void t6_pass_v(int rt, const double* tmpT, float* img, int P, int R, int lo, int hi,
                                          const double* s_kv, const int2* s_bv) {
  constexpr int kRows = 24;
  constexpr int kCg = kT6RThreads / kRows;      // column groups
  const int r = rt % kRows, cg = rt / kRows;
  const int yy = lo + r;
  if (yy > hi || cg >= kCg) return;
  const int2 bd = s_bv[yy];
  double k[KT];
#pragma unroll
  for (int y = 0; y < KT; ++y) k[y] = y < bd.y ? s_kv[y * R + yy] : 0.0;      // [tap][row]: lanes read neighbours
  float* orow = img + yy * P;
  const double* in = tmpT + bd.x;
  for (int xx = cg; xx < R; xx += kT6Chains * kCg) {
    const double* ip[kT6Chains];
    double sum[kT6Chains];
#pragma unroll
    for (int c = 0; c < kT6Chains; ++c) {
      ip[c] = in + (xx + c * kCg < R ? xx + c * kCg : xx) * kT6HP;
      sum[c] = 0.0;
    }
#pragma unroll
    for (int y = 0; y < KT; ++y)
      if (y < bd.y) {
#pragma unroll
        for (int c = 0; c < kT6Chains; ++c) sum[c] += ip[c][y] * k[y];
      }
#pragma unroll
    for (int c = 0; c < kT6Chains; ++c)
      if (xx + c * kCg < R) orow[xx + c * kCg] = static_cast<float>(sum[c]);
  }
}

// relu / leaky-relu on two fp32 values, packed to bf16x2 (low half = a)
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
__global__ void __launch_bounds__(kT6Threads, 1) k6_tower(const TowerParams p, const __grid_constant__ CUtensorMap map_out) {
  using L = T6Smem<C1, FUSE2>;
  constexpr int R = L::R;
  constexpr int H1 = R / 2;               // layer-1 output size
  constexpr int C2 = L::C2;
  constexpr int P = t6_img_pitch(R);
  constexpr int NB = L::NB;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* src = reinterpret_cast<float*>(smem + L::src);
  double* tmpd = reinterpret_cast<double*>(smem + L::tmpd);
  float* img = reinterpret_cast<float*>(smem + L::img);
  constexpr int NIMG = L::NIMG;
  float* img_alt = reinterpret_cast<float*>(smem + L::a1);            // FUSE2: second image buffer
  unsigned char* a1 = smem + L::a1;
  unsigned char* w1s = smem + L::w1;
  unsigned char* w2s = smem + L::w2;
  unsigned char* strip = smem + L::strip;
  unsigned char* stg2 = smem + L::stg2;
  double* s_kh = reinterpret_cast<double*>(smem + L::tabs);            // [12][R] (tap-major)
  double* s_kv = s_kh + R * 12;                                         // [8][R]
  int2* s_bh = reinterpret_cast<int2*>(s_kv + R * 8);
  int2* s_bv = s_bh + R;
  float* s_b2 = reinterpret_cast<float*>(smem + L::bias);
  uint64_t* mbar1 = reinterpret_cast<uint64_t*>(smem + L::bars);   // [4] layer-1 tile of group g done
  uint64_t* mbar2 = mbar1 + 4;                                      // layer-2 strip done
  uint64_t* full = mbar2 + 1;                                       // [NIMG][NB] image row block written (R -> T)
  uint64_t* empty = full + NIMG * NB;                               // [NIMG][NB] image row block consumed (T -> R)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + NB);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // which branch, which scans
  int br = 0, first = blockIdx.x;
  if (first >= p.ctas[0]) { first -= p.ctas[0]; br = 1; }
  if (br == 1 && first >= p.ctas[1]) { first -= p.ctas[1]; br = 2; }
  const int n_mine = p.ctas[br];
  const int H = p.rz.ph[br], W = p.rz.pw[br];
  const int ksh = p.rz.ksh[br], ksv = p.rz.ksv[br];

  // ---- one-time setup: zero image border, weights into operand layouts, tables, barriers, TMEM
  for (int e = tid; e < (R + 1) * P; e += kT6Threads) img[e] = 0.f;      // (the second buffer is zeroed with the im2col region)
  for (int e = tid; e < R * kT6HP; e += kT6Threads) tmpd[e] = 0.0;
  for (int e = tid; e < 4 * kT6A1Bytes / 16; e += kT6Threads) reinterpret_cast<uint4*>(a1)[e] = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < C1 * 4; e += kT6Threads) {          // 16-byte chunk (n, kc) of the layer-1 kernel
    const int n = e >> 2, kc = e & 3;
    *reinterpret_cast<uint4*>(w1s + (n >> 3) * 512 + kc * 128 + (n & 7) * 16) =
        *reinterpret_cast<const uint4*>(p.w1 + (static_cast<size_t>(br) * C1 + n) * kT6K1 + kc * 8);
  }
  {
    const double* kh = p.rz.kh[br];
    const double* kv = p.rz.kv[br];
    for (int e = tid; e < R * 12; e += kT6Threads) { const int x = e / R, xx = e - x * R; s_kh[e] = x < ksh ? kh[xx * ksh + x] : 0.0; }
    for (int e = tid; e < R * 8; e += kT6Threads) { const int y = e / R, yy = e - y * R; s_kv[e] = y < ksv ? kv[yy * ksv + y] : 0.0; }
    for (int e = tid; e < R; e += kT6Threads) { s_bh[e] = p.rz.bh[br][e]; s_bv[e] = p.rz.bv[br][e]; }
  }
  if (FUSE2) {
    for (int e = tid; e < 9 * C2 * 8; e += kT6Threads) {    // chunk c of row n of tap t
      const int c = e & 7, n = (e >> 3) % C2, t = e / (8 * C2);
      *reinterpret_cast<uint4*>(w2s + t * (C2 * 128) + n * 128 + ((c ^ (n & 7)) << 4)) =
          *reinterpret_cast<const uint4*>(p.w2 + (static_cast<size_t>(br) * C2 + n) * 576 + t * 64 + c * 8);
    }
    for (int e = tid; e < C2; e += kT6Threads) s_b2[e] = p.b2[br * C2 + e];
    // the planes start as zeros: pixel slot 20 of every even-column plane row is layer 2's zero pad
    for (int e = tid; e < L::strip_slots * 32; e += kT6Threads) reinterpret_cast<uint32_t*>(strip)[e] = 0u;
  }
  if (tid == 0) {
    for (int g = 0; g < 4; ++g) mbar_init(&mbar1[g], 1);
    mbar_init(mbar2, 1);
    for (int k = 0; k < NIMG * NB; ++k) {
      mbar_init(&full[k], kT6RThreads / 32);      // one arrival per R warp
      mbar_init(&empty[k], kT6TThreads / 32 - (FUSE2 ? 1 : 0));   // one arrival per T warp (FUSE2: warp 15 only issues UMMAs)
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // RML_T6_DBG=512: the long mbarrier waits in their "one warp polls, the others park on a bar.sync" form.
  // Same protocol, 2 % slower; it is the form compute-sanitizer's racecheck can follow (it does not model an
  // mbarrier acquire as a synchronisation edge), so the race-freedom of the hand-over is checked in this mode.
  const bool bar_form = (p.dbg & 512) != 0;

  // image rows of block k: what T round k needs beyond block k-1 (rounds overlap by the 3x3 halo)
  auto blk_lo = [](int k) { return k == 0 ? 0 : (FUSE2 ? 4 * kT6TH * k + 3 : 16 * k + 1); };
  auto blk_hi = [](int k) { const int h = FUSE2 ? 4 * kT6TH * k + 22 : 16 * k + 16; return h < R - 1 ? h : R - 1; };

  constexpr int kSlot01 = L::plane0_slots, kSlot10 = 2 * L::plane0_slots, kSlot11 = 2 * L::plane0_slots + L::plane1_slots;
  if (warp >= kT6TThreads / 32) {
    // =================================================================== R role: resize producer
    const int rt = tid - kT6TThreads;
    float* img_main = img;
    const bool hq = (p.dbg & 64) != 0;         // strided-column horizontal pass: fewer wavefronts, measured slower
    auto prefetch = [&](int64_t b) {
      const float* g = p.rz.feats + b * p.rz.F + p.rz.poff[br];       // 8-byte aligned rows
      for (int e = rt; e < (H * W) / 2; e += kT6RThreads) cp_async8(src + 2 * e, g + 2 * e);
    };
    if (first < p.B && !p.images) prefetch(first);
    uint32_t it = 0;
    for (int64_t b = first; b < p.B; b += n_mine, ++it) {
      if (!p.images) {
        cp_async_wait_all();
        bar_group(6, kT6RThreads);                               // projection landed; tmpd free (V pass of b-1 done)
        if (p.dbg & 2) {
        } else if (ksh <= 5) t6_pass_h<5>(rt, hq, src, tmpd, H, W, R, s_kh, s_bh);
        else if (ksh <= 8) t6_pass_h<8>(rt, hq, src, tmpd, H, W, R, s_kh, s_bh);
        else t6_pass_h<12>(rt, hq, src, tmpd, H, W, R, s_kh, s_bh);
        bar_group(6, kT6RThreads);
        if (b + n_mine < p.B) prefetch(b + n_mine);              // src is free from here on
      }
      const uint32_t ibuf = it % NIMG, iuse = it / NIMG;      // image buffer and how often it was used before
      float* img = ibuf ? img_alt : img_main;
      for (int k = 0; k < NB; ++k) {
        if (iuse > 0) {
          // the T role is done with this block of the image that was here before (two image buffers:
          // normally long ago)
          if (bar_form) {
            if (rt < 32) mbar_wait_relaxed(&empty[ibuf * NB + k], (iuse - 1) & 1);
            bar_group(6, kT6RThreads);
          } else {
            mbar_wait_relaxed(&empty[ibuf * NB + k], (iuse - 1) & 1);
          }
        }
        const int lo = blk_lo(k), hi = blk_hi(k);
        if (p.images) {
          const float* g = p.images + (static_cast<int64_t>(b) * 3 + br) * R * R;
          for (int e = lo * R + rt; e < (hi + 1) * R; e += kT6RThreads) img[(e / R) * P + (e % R)] = g[e];
        } else if (p.dbg & 2) {
        } else if (ksv <= 5) {
          t6_pass_v<5>(rt, tmpd, img, P, R, lo, hi, s_kv, s_bv);  // TF 'same' pads after: the image starts at (0,0)
        } else {
          t6_pass_v<8>(rt, tmpd, img, P, R, lo, hi, s_kv, s_bv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[ibuf * NB + k]);         // release: the rows are visible to the T role
      }
    }
  } else {
    // =================================================================== T role: layers 1 (+ 2)
    const int grp = warp >> 2;               // tile group 0..3
    const int q = warp & 3;                  // TMEM lane quarter of this warp
    const int m = q * 32 + lane;             // row of the group's tile
    const uint32_t tmem1 = tmem_base + grp * C1;                // layer-1 accumulator of this group (C1 <= 128)
    const uint32_t tmem2 = tmem_base + (FUSE2 ? 4 * C1 : 0);    // layer-2 accumulator (FUSE2: C1 = 64 -> column 256)
    const uint32_t idesc1 = umma_idesc(kCF32, kFmtF16, kFmtF16, 128, C1);
    uint32_t ph1 = 0;                        // mbarrier phase of this group's layer-1 tile
    uint32_t gstrip = 0;                     // running strip counter (FUSE2): phase of mbar2
    const uint32_t idesc2 = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, C2);
    // layer-2 epilogue of one finished strip (T group 0): TMEM -> bias + act -> bf16 -> global
    auto epilogue2 = [&](int64_t eb, int es, uint32_t acc) {
      tc_fence_after();
      const int oyl = m / kT6Pitch, ox = m - oyl * kT6Pitch;
      const int oy = kT6TH * es + oyl;
      const uint32_t taddr = acc + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t v[32];
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (oyl < kT6TH && ox < H1 / 2) {
        __nv_bfloat16* o = p.out + (((eb * 3 + br) * (H1 / 2) + oy) * (H1 / 2) + ox) * C2;
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
    };
    constexpr int kRowsPerStrip = FUSE2 ? 2 * kT6TH + 1 : 8;             // layer-1 rows per round of 4 tiles
    constexpr int kPixPerStrip = kRowsPerStrip * H1;                     // 440 (dnn) / 512 (sgan)
    static_assert(kPixPerStrip <= 512, "four 128-pixel tiles per round");
    if constexpr (FUSE2) {
      // ---- dnn: software-pipelined rounds.  A round = one strip of 11 layer-1 rows (4 tiles) and the
      // layer-2 strip they feed.  Iteration r: [barrier] -> one thread issues the layer-1 GEMMs of
      // round r and THEN the 36 layer-2 UMMAs of round r-1 (the tensor pipe runs in issue order, so the
      // short layer-1 GEMMs must not queue behind them) -> every warp converts its layer-1 tile to
      // bf16 in registers and builds the im2col tile of round r+1 while the layer-2 UMMAs run -> only
      // the 128-bit stores of the strip wait for them (the planes are single-buffered: 62 KB).
      // The layer-2 accumulator is double-buffered so that group 0 drains strip r-2 under the UMMAs.
      const int64_t n_img = first < p.B ? (p.B - first + n_mine - 1) / n_mine : 0;
      const uint32_t n_rounds = static_cast<uint32_t>(n_img) * NB;
      const bool warp_dead = grp * 128 + q * 32 >= kPixPerStrip;   // 440 of 512 pixel slots are used
      // the last warp holds no pixel: it becomes the UMMA issuer.  (Issued by warp 0 next to its pixel
      // work, the 44 UMMAs of a round — one thread, ~2.5 k cycles — made group 0 the straggler at every
      // barrier: half of the tile role's time was barrier wait.)
      const bool mma_warp = warp == kT6TThreads / 32 - 1;
      const int gthreads = grp == 3 ? 96 : 128;                    // group barrier without the issuer
      const int pidx = grp * 128 + m;                               // pixel of this thread inside a round
      const int prow = pidx / H1, x1 = pidx - prow * H1;
      const bool store = pidx < kPixPerStrip;
      const int slot = ((prow & 1) ? ((x1 & 1) ? kSlot11 : kSlot10) : ((x1 & 1) ? kSlot01 : 0)) +
                       (prow >> 1) * kT6Pitch + (x1 >> 1);
      unsigned char* dst_row = strip + slot * 128;
      const uint32_t sw = slot & 7;             // the swizzle follows the absolute address (strip is 1024-aligned)
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t tmemA0 = tmem_base + 4 * C1 + 64;              // 2 x four A tiles of 16 columns (double-buffered)
      const bool early = !(p.dbg & 32);         // gather round r+1 BEFORE awaiting the layer-1 commit of round r
      auto im2col = [&](uint32_t r) {
        const uint32_t im = r / NB;
        const int s = static_cast<int>(r - im * NB);
        // image rows of this round are in smem (acquire); one warp of the group polls, the rest park
        const uint32_t ibuf = im % NIMG, iuse = im / NIMG;
        // every warp acquires the barrier itself: it has normally completed long ago (the R role runs
        // ahead), and a group barrier behind one polling warp costs a round trip per round
        if (bar_form) {
          if (q == 0) mbar_wait_relaxed(&full[ibuf * NB + s], iuse & 1);
          bar_group(1 + grp, gthreads);
        } else {
          mbar_wait_relaxed(&full[ibuf * NB + s], iuse & 1);
        }
        if (!warp_dead && !(p.dbg & 8)) {
          const int y1 = 2 * kT6TH * s + prow;
          const bool live = store && y1 < H1;                       // y1 == H1 is layer 2's zero pad row
          float v[9];
          if (live) {
            const float* ip = (ibuf ? img_alt : img) + (2 * y1) * P + 2 * x1;
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
          // K order (fp16 pairs): hi0..hi8, ONE | lo0..lo8, 0 | hi0..hi8, ONE | 0, 0 against kernel rows
          // whi0..8, bias_hi | whi0..8, 0 | wlo0..8, bias_lo | 0, 0.  A dead pixel keeps an all-zero row
          // (ONE = 0), so its output is exactly 0 = layer 2's pad value.
          uint32_t hw[5], lw[5];
          const float one = live ? 1.f : 0.f;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float a0 = v[2 * i], a1v = i < 4 ? v[2 * i + 1] : one;
            const __half2 h2 = __floats2half2_rn(a0, a1v);
            const float r0 = a0 - __low2float(h2), r1 = i < 4 ? a1v - __high2float(h2) : 0.f;
            const __half2 l2 = __floats2half2_rn(r0, r1);
            hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
            lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          // straight to tensor memory (lane = pixel, 16 columns = K 32): the layer-1 A operand costs no
          // shared-memory store and no shared-memory read
          const uint32_t av[16] = {hw[0], hw[1], hw[2], hw[3], hw[4], lw[0], lw[1], lw[2],
                                   lw[3], lw[4], hw[0], hw[1], hw[2], hw[3], hw[4], 0u};
          tmem_st_32x16(tmemA0 + (r & 1) * 64 + grp * 16 + lane_off, av);
          tmem_st_wait();
        }
        // this warp no longer reads the previous block of image rows (and, in the last round, this one)
        __syncwarp();
        if (lane == 0) {
          if (s > 0) mbar_arrive(&empty[ibuf * NB + s - 1]);
          if (s == NB - 1) mbar_arrive(&empty[ibuf * NB + s]);
        }
        tc_fence_before();
      };
      uint32_t pk[32];
      if (n_rounds && !mma_warp) im2col(0);
      for (uint32_t r = 0; r <= n_rounds; ++r) {
        bar_group(5, kT6TThreads);        // im2col tiles of round r in tensor memory, strip r-1 in shared memory
        if (mma_warp) {
          tc_fence_after();
          if (elect_one()) {
            if (r < n_rounds && !(p.dbg & 16)) {
              const uint32_t b_addr = smem_u32(w1s);
#pragma unroll
              for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  umma_f16_ts(tmem_base + g * C1, tmemA0 + (r & 1) * 64 + g * 16 + 8 * k, t6_desc_noswz(b_addr + k * 256), idesc1, k != 0);
            }
            if (r < n_rounds) umma_commit(&mbar1[0]);
            if (r >= 1) {
              const uint32_t sbase = smem_u32(strip);
              const uint32_t wbase = smem_u32(w2s);
              const uint32_t acc = tmem2 + ((r - 1) & 1) * 32;
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                if (p.dbg & 1) break;
                const int kh3 = t / 3, kw3 = t - 3 * kh3;
                const int slot0 = ((kh3 & 1) ? ((kw3 & 1) ? kSlot11 : kSlot10) : ((kw3 & 1) ? kSlot01 : 0)) +
                                  (kh3 >> 1) * kT6Pitch + (kw3 >> 1);
                const uint64_t da = umma_desc_k_sw128(sbase + slot0 * 128);
                const uint64_t db = umma_desc_k_sw128(wbase + t * (C2 * 128));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_f16(acc, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc2, (t | ks) != 0);
              }
              umma_commit(mbar2);
            }
          }
          __syncwarp();
          continue;
        }
        // the im2col tiles are double-buffered in tensor memory, so the gather of round r+1 does not wait
        // for the layer-1 UMMAs of round r: it hides their commit -> mbarrier round trip
        if (early && r + 1 < n_rounds) im2col(r + 1);
        // group 0 drains strip r-2 (its UMMAs were awaited by every warp in iteration r-1)
        if (grp == 0 && r >= 2) {
          const uint32_t rr = r - 2, im = rr / NB;
          epilogue2(first + static_cast<int64_t>(im) * n_mine, static_cast<int>(rr - im * NB), tmem2 + (rr & 1) * 32);
        }
        if (r < n_rounds) {
          mbar_wait(&mbar1[0], r & 1);
          tc_fence_after();
          if (!warp_dead && !(p.dbg & 4)) {
#pragma unroll
            for (int c0 = 0; c0 < C1; c0 += 32) {
              uint32_t v[32];
              tmem_ld_32x32(tmem1 + lane_off + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; e += 2)
                pk[(c0 + e) >> 1] = act_pack<ACT>(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), p.alpha);
            }
          }
          tc_fence_before();
        }
        if (!early && r + 1 < n_rounds) im2col(r + 1);
        if (r >= 1) {
          // the strip planes are still being read by the layer-2 UMMAs of round r-1
          if (bar_form) {
            if (q == 0) mbar_wait_relaxed(mbar2, (r - 1) & 1);
            bar_group(1 + grp, gthreads);
          } else {
            mbar_wait_relaxed(mbar2, (r - 1) & 1);
          }
        }
        if (r < n_rounds) {
          if (store && !warp_dead && !(p.dbg & 4)) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(dst_row + ((c ^ sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          }
          fence_proxy_async_smem();
        }
      }
      if (grp == 0 && n_rounds >= 1) {
        const uint32_t rr = n_rounds - 1, im = rr / NB;
        tc_fence_after();
        epilogue2(first + static_cast<int64_t>(im) * n_mine, static_cast<int>(rr - im * NB), tmem2 + (rr & 1) * 32);
      }
    } else {
    uint32_t it = 0;
    for (int64_t b = first; b < p.B; b += n_mine, ++it) {
      for (int s = 0; s < NB; ++s) {
        const int y1_0 = FUSE2 ? 2 * kT6TH * s : 8 * s;       // first layer-1 row of the round
        const int pidx = grp * 128 + m;                        // pixel of this thread inside the round
        const int r = pidx / H1, x1 = pidx - r * H1;
        const int y1 = y1_0 + r;
        const bool live = pidx < kPixPerStrip && y1 < H1;      // y1 == H1 is layer 2's zero pad row
        // FUSE2: the last warps of group 3 hold no pixel of any round (440 of 512): their im2col rows
        // stay the zeros written at setup and they skip the gather and the epilogue
        const bool warp_dead = FUSE2 && grp * 128 + q * 32 >= kPixPerStrip;
        // image rows of this round are in smem (acquire); one warp of the group polls, the rest park
        if (q == 0) mbar_wait_relaxed(&full[s], it & 1);
        bar_group(1 + grp, 128);
        // -- im2col row of this pixel: 9 taps, fp16 hi / lo split
        if (!(p.dbg & 8) && !warp_dead) {
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
          // K order (fp16 pairs, one cvt.rn.f16x2 each): hi0..hi8, ONE | lo0..lo8, 0 | hi0..hi8, ONE | 0, 0
          // against kernel rows whi0..8, bias_hi | whi0..8, 0 | wlo0..8, bias_lo | 0, 0.  A dead pixel
          // keeps an all-zero row (ONE = 0), so its output is exactly 0 = layer 2's pad value.
          uint32_t hw[5], lw[5];
          const float one = live ? 1.f : 0.f;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float a0 = v[2 * i], a1v = i < 4 ? v[2 * i + 1] : one;
            const __half2 h2 = __floats2half2_rn(a0, a1v);
            const float r0 = a0 - __low2float(h2), r1 = i < 4 ? a1v - __high2float(h2) : 0.f;
            const __half2 l2 = __floats2half2_rn(r0, r1);
            hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
            lw[i] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          unsigned char* row = a1 + grp * kT6A1Bytes + (m >> 3) * 512 + (m & 7) * 16;
          *reinterpret_cast<uint4*>(row) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(row + 128) = make_uint4(hw[4], lw[0], lw[1], lw[2]);
          *reinterpret_cast<uint4*>(row + 256) = make_uint4(lw[3], lw[4], hw[0], hw[1]);
          *reinterpret_cast<uint4*>(row + 384) = make_uint4(hw[2], hw[3], hw[4], 0u);
        }
        // this warp no longer reads the previous block of image rows (and, in the last round, this one)
        __syncwarp();
        if (lane == 0) {
          if (s > 0) mbar_arrive(&empty[s - 1]);
          if (s == NB - 1) mbar_arrive(&empty[s]);
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
              if (!(p.dbg & 16)) umma_f16(tmem1, t6_desc_noswz(a_addr + k * 256), t6_desc_noswz(b_addr + k * 256), idesc1, k != 0);
            umma_commit(&mbar1[grp]);
          }
          __syncwarp();
        }
        mbar_wait(&mbar1[grp], ph1);
        ph1 ^= 1;
        tc_fence_after();
        if (FUSE2 && gstrip > 0) {
          // the strip planes are still being read by the layer-2 MMAs of the previous strip
          if (q == 0) mbar_wait_relaxed(mbar2, (gstrip - 1) & 1);
          bar_group(1 + grp, 128);
        }
        // -- layer-1 epilogue: activation -> bf16 -> strip planes (FUSE2) or global NHWC
        if (!(p.dbg & 4) && !warp_dead) {
          const uint32_t taddr = tmem1 + (static_cast<uint32_t>(q * 32) << 16);
          unsigned char* dst_row = nullptr;
          uint32_t sw = 0;
          if (FUSE2) {
            const int py = r & 1, a = r >> 1, px = x1 & 1, j = x1 >> 1;
            const int slot = (py ? (px ? kSlot11 : kSlot10) : (px ? kSlot01 : 0)) + a * kT6Pitch + j;
            dst_row = strip + slot * 128;
            sw = slot & 7;                      // the swizzle follows the absolute address (strip is 1024-aligned)
          }
          const bool store = pidx < kPixPerStrip && (FUSE2 || y1 < H1);
          // !FUSE2: a thread holds one pixel's channels, so a direct store would touch 32 half-filled
          // sectors per instruction (measured: 141 k cycles per image, the whole kernel).  The warp
          // transposes each 32-pixel x 64-byte block through its quarter of the group's im2col tile
          // (free once the layer-1 MMAs have retired) and writes 64 contiguous bytes per 4 lanes.
          unsigned char* stg = a1 + grp * kT6A1Bytes + q * 2048;
          __nv_bfloat16* gtile = p.out + ((static_cast<int64_t>(b) * 3 + br) * H1 * H1 +
                                          static_cast<int64_t>(y1_0) * H1 + grp * 128 + q * 32) * C1;
          // !FUSE2 (default): the warp's 32 pixels x 64 channels are staged as 128-byte rows under the 128B
          // swizzle (conflict-free 128-bit stores) and leave as ONE tiled TMA store of full 128-byte lines
          // (box {64 ch, 32 px} of the [pixels][C1] tensor) — no LDS / STG chain on the load/store path the
          // resize role shares.  (RML_T6_DBG=1024 keeps the warp-transpose + STG.128 form for comparison.)
          const bool tma_out = !FUSE2 && !(p.dbg & 1024);
          unsigned char* stg4 = q < 2 ? a1 + grp * kT6A1Bytes + q * 4096 : stg2 + grp * 8192 + (q - 2) * 4096;
          const int32_t pix0 = static_cast<int32_t>((static_cast<int64_t>(b) * 3 + br) * H1 * H1 + y1_0 * H1 + grp * 128 + q * 32);
#pragma unroll
          for (int c0 = 0; c0 < C1; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c0, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2)
              pk[e >> 1] = act_pack<ACT>(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), p.alpha);
            if (tma_out) {
              if ((c0 & 32) == 0) {
                // the previous store out of this staging buffer has read it (waited after the TMEM load
                // and the conversion, which do not need the buffer)
                if (lane == 0) bulk_wait_read<0>();
                __syncwarp();
              }
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(stg4 + lane * 128 + (((((c0 & 32) >> 3) + c) ^ (lane & 7)) << 4)) =
                    make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
              if (c0 & 32) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&map_out, stg4, c0 - 32, pix0);
                  bulk_commit();
                }
              }
            } else if (FUSE2) {
              if (store) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  *reinterpret_cast<uint4*>(dst_row + ((((c0 >> 3) + c) ^ sw) << 4)) =
                      make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(stg + lane * 64 + (((c ^ (lane >> 1)) & 3) << 4)) =
                    make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int px = 8 * i + (lane >> 2), c = lane & 3;
                const uint4 w = *reinterpret_cast<const uint4*>(stg + px * 64 + (((c ^ (px >> 1)) & 3) << 4));
                const int pp = grp * 128 + q * 32 + px;                  // pixel inside the round
                if (pp < kPixPerStrip && y1_0 + pp / H1 < H1)
                  *reinterpret_cast<uint4*>(gtile + static_cast<int64_t>(px) * C1 + c0 + 8 * c) = w;
              }
              __syncwarp();
            }
          }
        }
        tc_fence_before();
        if (!FUSE2) {
          // warps 0-1 staged in the group's im2col tile: their TMA stores must have read it before the
          // next round's gather overwrites it
          if (q < 2 && lane == 0) bulk_wait_read<0>();
          bar_group(1 + grp, 128);          // the tile's TMEM and im2col buffer are free for the next round
          continue;
        }
        fence_proxy_async_smem();
        bar_group(5, kT6TThreads);          // the whole strip is in shared memory
        // -- layer 2: 9 taps x 4 UMMAs (K = 64 channels) straight from the parity planes.  Shared-memory
        // bandwidth bounds this kernel (ncu: LSU + tensor wavefronts = 94 % of the data pipe): a
        // 128 x 32 x 16 UMMA reads its whole 4 KB A slice for 32 columns of work, 147 KB per strip.
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sbase = smem_u32(strip);
            const uint32_t wbase = smem_u32(w2s);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (p.dbg & 1) break;
              const int kh3 = t / 3, kw3 = t - 3 * kh3;
              const int slot0 = ((kh3 & 1) ? ((kw3 & 1) ? kSlot11 : kSlot10) : ((kw3 & 1) ? kSlot01 : 0)) +
                                (kh3 >> 1) * kT6Pitch + (kw3 >> 1);
              const uint64_t da = umma_desc_k_sw128(sbase + slot0 * 128);
              const uint64_t db = umma_desc_k_sw128(wbase + t * (C2 * 128));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_f16(tmem2, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc2, (t | ks) != 0);
            }
            umma_commit(mbar2);
          }
          __syncwarp();
        }
        ++gstrip;
        if (grp == 0) {
          if (q == 0) mbar_wait_relaxed(mbar2, (gstrip - 1) & 1);
          bar_group(1, 128);
          epilogue2(b, s, tmem2);
        }
      }
    }
    // drain: the last strip's MMAs must retire before TMEM is released (group 0 has waited already)
    if (FUSE2 && gstrip > 0 && grp != 0) mbar_wait(mbar2, (gstrip - 1) & 1);
    if (!FUSE2 && lane == 0) bulk_wait_all<0>();      // this warp's TMA stores have left shared memory and landed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace rml
