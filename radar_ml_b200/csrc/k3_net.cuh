// K3/K4/K5 — the dnn.py / sgan.py classifier forward pass.
//   k3_resize_pil   dnn.py:240-245 / sgan.py:676-681: PIL Image.resize((R,R), BICUBIC) of the
//                   three projections as a fixed separable operator (Pillow Resample.c tables
//                   computed on the host), double accumulation, float32 store — bit-exact.
//   k4_conv3x3s2    Keras Conv2D(3x3, strides 2, 'same') + folded BatchNorm + ReLU/LeakyReLU
//                   (dnn.py:45-52, sgan.py:132-154), NHWC, fp32 CUDA-core direct conv
//                   (kept for shapes the implicit GEMM cannot take and for RML_IGEMM=0);
//   k4_conv_igemm   the same layer for Cin % 64 == 0 as a tcgen05 implicit GEMM (strided 4-D TMA).
//   k5_dense_stack  Flatten -> Dense 64 -> Dense 64 -> Dense C -> softmax / Z/(Z+1)
//                   (dnn.py:78-85, sgan.py:185-213).  The K = 38 400 / 24 576 contraction runs
//                   on tcgen05 (kind::f16, bf16 operands, fp32 accumulate in TMEM, TMA-fed
//                   128B-swizzled tiles); the two 64-wide layers and the head are fused into
//                   the TMEM epilogue in fp32.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace rml {

// ------------------------------------------------------------------------------ K3 resize
struct ResizeParams {
  const float* feats;     // [B][F] projections already scaled (p-127.5)/127.5 by K1
  float* images;          // [B][3][R][R]
  int64_t B;
  int F, R;
  int ph[3], pw[3], poff[3];   // projection heights / widths / offsets inside a feature row
  const double* kh[3];    // horizontal tables [R][ksh]   (pw -> R)
  const double* kv[3];    // vertical tables   [R][ksv]   (ph -> R)
  const int2* bh[3];      // [R] (first input index, taps)
  const int2* bv[3];
  int ksh[3], ksv[3];
};

// The two passes of Pillow's ImagingResample for one image held in shared memory, by one CTA.
// Arithmetic is Pillow's: for each output, ss = 0.0; ss += (double)in[xmin + x] * k[x] for x in
// [0, xmax) in that order; store (float)ss — what keeps the result bit-identical to PIL.
// Mapping: a thread owns one output column and walks over rows, so the horizontal coefficients
// of its column sit in registers (no per-tap coefficient load, no index division) and the
// vertical coefficients of a row are warp-uniform loads.  The tap loop is unrolled to 5, 8 or 12
// slots (5 taps when enlarging, 7 for 176 -> 128, 11 for 176 -> 80): a predicated-off slot still
// issues, so the unroll follows the table; longer tables take the plain loop.
constexpr int kResizeTaps = 12;

template <int KT>
__device__ __forceinline__ void resize_rows_h(const float* src, float* tmp, int H, int W, int R,
                                              const double* kh, const int2* bh, int ksh) {
  const int groups = blockDim.x / R;
  const int xx = threadIdx.x % R, grp = threadIdx.x / R;
  if (grp >= groups) return;
  const int2 bd = bh[xx];
  double k[KT];
#pragma unroll
  for (int x = 0; x < KT; ++x) k[x] = x < bd.y ? kh[xx * ksh + x] : 0.0;
  for (int y = grp; y < H; y += groups) {
    const float* in = src + y * W + bd.x;
    double ss = 0.0;
#pragma unroll
    for (int x = 0; x < KT; ++x)
      if (x < bd.y) ss += static_cast<double>(in[x]) * k[x];
    tmp[y * R + xx] = static_cast<float>(ss);
  }
}

template <int KT>
__device__ __forceinline__ void resize_rows_v(const float* tmp, float* out, int out_pitch, int R,
                                              const double* kv, const int2* bv, int ksv) {
  const int groups = blockDim.x / R;
  const int xx = threadIdx.x % R, grp = threadIdx.x / R;
  if (grp >= groups) return;
  for (int yy = grp; yy < R; yy += groups) {
    const int2 bd = bv[yy];
    const float* in = tmp + bd.x * R + xx;
    const double* k = kv + yy * ksv;
    double ss = 0.0;
#pragma unroll
    for (int y = 0; y < KT; ++y)
      if (y < bd.y) ss += static_cast<double>(in[y * R]) * k[y];
    out[yy * out_pitch + xx] = static_cast<float>(ss);
  }
}

__device__ __forceinline__ void resize_pass_h(const float* src, float* tmp, int H, int W, int R,
                                              const double* kh, const int2* bh, int ksh) {
  if (ksh <= kResizeTaps && R <= static_cast<int>(blockDim.x)) {
    if (ksh <= 5) resize_rows_h<5>(src, tmp, H, W, R, kh, bh, ksh);
    else if (ksh <= 8) resize_rows_h<8>(src, tmp, H, W, R, kh, bh, ksh);
    else resize_rows_h<kResizeTaps>(src, tmp, H, W, R, kh, bh, ksh);
  } else {
    for (int e = threadIdx.x; e < H * R; e += blockDim.x) {
      const int y = e / R, xx = e - y * R;
      const int2 bd = bh[xx];
      double ss = 0.0;
      for (int x = 0; x < bd.y; ++x) ss += static_cast<double>(src[y * W + bd.x + x]) * kh[xx * ksh + x];
      tmp[e] = static_cast<float>(ss);
    }
  }
}

// out[yy * out_pitch + xx] for yy, xx < R (out may be shared or global memory)
__device__ __forceinline__ void resize_pass_v(const float* tmp, float* out, int out_pitch, int R,
                                              const double* kv, const int2* bv, int ksv) {
  if (ksv <= kResizeTaps && R <= static_cast<int>(blockDim.x)) {
    if (ksv <= 5) resize_rows_v<5>(tmp, out, out_pitch, R, kv, bv, ksv);
    else if (ksv <= 8) resize_rows_v<8>(tmp, out, out_pitch, R, kv, bv, ksv);
    else resize_rows_v<kResizeTaps>(tmp, out, out_pitch, R, kv, bv, ksv);
  } else {
    for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
      const int yy = e / R, xx = e - yy * R;
      const int2 bd = bv[yy];
      double ss = 0.0;
      for (int y = 0; y < bd.y; ++y) ss += static_cast<double>(tmp[(bd.x + y) * R + xx]) * kv[yy * ksv + y];
      out[yy * out_pitch + xx] = static_cast<float>(ss);
    }
  }
}

// one CTA per (scan, branch); smem: projection + horizontal-pass result
__global__ void __launch_bounds__(256) k3_resize_pil(const ResizeParams p) {
  extern __shared__ float rs_smem[];
  const int br = blockIdx.y;
  const int H = p.ph[br], W = p.pw[br], R = p.R;
  float* src = rs_smem;             // [H][W]
  float* tmp = rs_smem + H * W;     // [H][R]
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* g = p.feats + b * p.F + p.poff[br];
    for (int e = threadIdx.x; e < H * W; e += blockDim.x) src[e] = g[e];
    __syncthreads();
    // horizontal pass (ImagingResampleHorizontal_32bpc), then vertical: double accumulate, float store
    resize_pass_h(src, tmp, H, W, R, p.kh[br], p.bh[br], p.ksh[br]);
    __syncthreads();
    resize_pass_v(tmp, p.images + (b * 3 + br) * static_cast<int64_t>(R) * R, R, R, p.kv[br], p.bv[br],
                  p.ksv[br]);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ K4 conv
constexpr int kConvCoutTile = 32;
struct ConvParams {
  const float* in;        // [n_img][H][W][Cin] fp32 (n_img = scans * 3, branch = img % 3)
  void* out;              // [n_img][Ho][Wo][Cout] fp32 or bf16
  const float* w[3];      // per branch [3][3][Cin][Cout] (Keras HWIO, BN folded)
  const float* bias[3];   // per branch [Cout]
  int64_t n_img;
  int H, W, Cin, Cout, Ho, Wo, pad_t, pad_l;
  int act;                // 0 none, 1 relu, 2 leaky relu
  float alpha;
  int out_bf16;
};

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v >= 0.f ? v : alpha * v;
  return v;
}

// grid: (pixel blocks, Cout/32, 3 branches); each thread = one output pixel x 32 channels;
// the 9*Cin*32 weights of this (branch, channel group) live in shared memory.
__global__ void __launch_bounds__(256) k4_conv3x3s2(const ConvParams p) {
  extern __shared__ float cw[];   // [9][Cin][32]
  const int br = blockIdx.z;
  const int cg = blockIdx.y;
  const float* wsrc = p.w[br];
  const int n_w = 9 * p.Cin * kConvCoutTile;
  for (int e = threadIdx.x; e < n_w; e += blockDim.x) {
    const int co = e % kConvCoutTile;
    const int t = e / kConvCoutTile;          // tap * Cin + ci
    cw[e] = wsrc[t * p.Cout + cg * kConvCoutTile + co];
  }
  __syncthreads();
  const int64_t per_branch = (p.n_img / 3) * p.Ho * p.Wo;
  const float* bias = p.bias[br] + cg * kConvCoutTile;
  for (int64_t pix = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; pix < per_branch;
       pix += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(pix % p.Wo);
    const int oy = static_cast<int>((pix / p.Wo) % p.Ho);
    const int64_t scan = pix / (static_cast<int64_t>(p.Wo) * p.Ho);
    const int64_t img = scan * 3 + br;
    float acc[kConvCoutTile];
#pragma unroll
    for (int c = 0; c < kConvCoutTile; ++c) acc[c] = bias[c];
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * 2 + kh - p.pad_t;
      if (iy < 0 || iy >= p.H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * 2 + kw - p.pad_l;
        if (ix < 0 || ix >= p.W) continue;
        const float* xin = p.in + ((img * p.H + iy) * p.W + ix) * p.Cin;
        const float* wt = cw + (kh * 3 + kw) * p.Cin * kConvCoutTile;
        if ((p.Cin & 3) == 0) {
          for (int ci = 0; ci < p.Cin; ci += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(xin + ci);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4* w4 = reinterpret_cast<const float4*>(wt + (ci + u) * kConvCoutTile);
#pragma unroll
              for (int c4 = 0; c4 < kConvCoutTile / 4; ++c4) {
                const float4 wv = w4[c4];
                acc[c4 * 4 + 0] = fmaf(xs[u], wv.x, acc[c4 * 4 + 0]);
                acc[c4 * 4 + 1] = fmaf(xs[u], wv.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(xs[u], wv.z, acc[c4 * 4 + 2]);
                acc[c4 * 4 + 3] = fmaf(xs[u], wv.w, acc[c4 * 4 + 3]);
              }
            }
          }
        } else {
          for (int ci = 0; ci < p.Cin; ++ci) {
            const float x = xin[ci];
#pragma unroll
            for (int c = 0; c < kConvCoutTile; ++c) acc[c] = fmaf(x, wt[ci * kConvCoutTile + c], acc[c]);
          }
        }
      }
    }
    const int64_t o = ((img * p.Ho + oy) * p.Wo + ox) * p.Cout + cg * kConvCoutTile;
    if (p.out_bf16) {
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + o;
#pragma unroll
      for (int c = 0; c < kConvCoutTile; c += 2)
        *reinterpret_cast<__nv_bfloat162*>(out + c) =
            __floats2bfloat162_rn(apply_act(acc[c], p.act, p.alpha), apply_act(acc[c + 1], p.act, p.alpha));
    } else {
      float* out = reinterpret_cast<float*>(p.out) + o;
#pragma unroll
      for (int c = 0; c < kConvCoutTile; c += 4)
        *reinterpret_cast<float4*>(out + c) =
            make_float4(apply_act(acc[c], p.act, p.alpha), apply_act(acc[c + 1], p.act, p.alpha),
                        apply_act(acc[c + 2], p.act, p.alpha), apply_act(acc[c + 3], p.act, p.alpha));
    }
  }
}

// First tower layer (Cin = 1, K = 9): a lane owns CPL = Cout/32 output channels with its 9*CPL
// weights in registers and walks over pixels; the 9 input taps are warp-uniform loads and every
// pixel is written as one fully coalesced Cout*2-byte bf16 row.  Bound by the activation write.
struct Conv1Params {
  const float* in;        // [n_img][H][W] fp32 (single channel)
  __nv_bfloat16* out;     // [n_img][Ho][Wo][Cout]
  const float* w[3];      // per branch [9][Cout]
  const float* bias[3];
  int64_t n_img;
  int H, W, Cout, Ho, Wo, pad_t, pad_l;
  int act;
  float alpha;
};
constexpr int kConv1Run = 64;   // (host grid sizing) pixels per warp task, roughly

// warp task = one output row of one image; 4 consecutive output pixels per step share their
// 3 x 9 input taps (27 warp-uniform loads in flight instead of 9 dependent ones per pixel)
template <int CPL>
__global__ void __launch_bounds__(256) k4_conv1_cin1(const Conv1Params p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t tasks = p.n_img * p.Ho;
  int cur_br = -1;
  float w[9][CPL], bs[CPL];
  for (int64_t task = warp_global; task < tasks; task += n_warps) {
    const int64_t img = task / p.Ho;
    const int oy = static_cast<int>(task - img * p.Ho);
    const int br = static_cast<int>(img % 3);
    if (br != cur_br) {
      cur_br = br;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < CPL; ++c) w[t][c] = p.w[br][t * p.Cout + lane * CPL + c];
#pragma unroll
      for (int c = 0; c < CPL; ++c) bs[c] = p.bias[br][lane * CPL + c];
    }
    const float* in = p.in + img * static_cast<int64_t>(p.H) * p.W;
    __nv_bfloat16* out = p.out + (img * p.Ho + oy) * static_cast<int64_t>(p.Wo) * p.Cout + lane * CPL;
    const int iy0 = oy * 2 - p.pad_t;
    for (int ox0 = 0; ox0 < p.Wo; ox0 += 4) {
      const int ix0 = ox0 * 2 - p.pad_l;
      float x[3][9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int iy = iy0 + kh;
        const bool rowok = iy >= 0 && iy < p.H;
#pragma unroll
        for (int cx = 0; cx < 9; ++cx) {
          const int ix = ix0 + cx;
          x[kh][cx] = (rowok && ix >= 0 && ix < p.W) ? __ldg(in + iy * p.W + ix) : 0.f;
        }
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        if (ox0 + px < p.Wo) {
          float acc[CPL];
#pragma unroll
          for (int c = 0; c < CPL; ++c) acc[c] = bs[c];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
              for (int c = 0; c < CPL; ++c) acc[c] = fmaf(x[kh][2 * px + kw], w[kh * 3 + kw][c], acc[c]);
          __nv_bfloat16* o = out + static_cast<int64_t>(ox0 + px) * p.Cout;
#pragma unroll
          for (int c = 0; c < CPL; c += 2)
            *reinterpret_cast<__nv_bfloat162*>(o + c) =
                __floats2bfloat162_rn(apply_act(acc[c], p.act, p.alpha), apply_act(acc[c + 1], p.act, p.alpha));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ K3+K4 fused
// PIL resize (k3_resize_pil arithmetic, bit-exact) + the first tower layer (Cin = 1) in one CTA
// per (scan, branch): the projection, the horizontal-pass result and the resized R x R image
// never leave shared memory; only the bf16 NHWC activation of layer 1 is written.
struct ResizeConv1Params {
  ResizeParams rz;        // rz.images unused
  __nv_bfloat16* out;     // [n_img][Ho][Wo][Cout]
  const float* w[3];      // per branch [9][Cout]
  const float* bias[3];
  int Cout, Ho, Wo, pad_t, pad_l;
  int act;
  float alpha;
};

// Shared-memory layout (floats): projection [H][W] | horizontal pass [H][R] | image with a zero
// border [(R+2)][R+16], each region rounded up to 16 bytes.  The border makes every 3x3 tap of the
// first layer an unconditional load (TF 'same' padding is just the zeros around the image), and
// with pad_l == 0 the 9 taps of a row of four output pixels are two 16-byte loads and one scalar.
__host__ __device__ constexpr int k34_round4(int v) { return (v + 3) & ~3; }
__host__ __device__ constexpr int k34_pitch(int R) { return R + 16; }
__host__ __device__ constexpr int k34_smem_floats(int H, int W, int R) {
  return k34_round4(H * W) + k34_round4(H * R) + (R + 2) * k34_pitch(R);
}

template <int CPL>
__global__ void __launch_bounds__(256) k34_resize_conv1(const ResizeConv1Params p) {
  extern __shared__ __align__(16) float rc_smem[];
  const int br = blockIdx.y;
  const int H = p.rz.ph[br], W = p.rz.pw[br], R = p.rz.R;
  const int P = k34_pitch(R);
  float* src = rc_smem;                           // [H][W]
  float* tmp = src + k34_round4(H * W);           // [H][R]
  float* img = tmp + k34_round4(H * R);           // [(R+2)][P], interior at (pad_t, colpad)
  const int colpad = p.pad_l ? 4 : 0;
  const int coff = colpad - p.pad_l;              // column of tap ix = -pad_l
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float w[9][CPL], bs[CPL];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < CPL; ++c) w[t][c] = p.w[br][t * p.Cout + lane * CPL + c];
#pragma unroll
  for (int c = 0; c < CPL; ++c) bs[c] = p.bias[br][lane * CPL + c];
  const double* kh = p.rz.kh[br];
  const double* kv = p.rz.kv[br];
  const int2* bh = p.rz.bh[br];
  const int2* bv = p.rz.bv[br];
  const int ksh = p.rz.ksh[br], ksv = p.rz.ksv[br];
  for (int e = threadIdx.x; e < (R + 2) * P; e += blockDim.x) img[e] = 0.f;   // the border stays zero
  __syncthreads();
  for (int64_t b = blockIdx.x; b < p.rz.B; b += gridDim.x) {
    const float* g = p.rz.feats + b * p.rz.F + p.rz.poff[br];
    for (int e = threadIdx.x; e < H * W; e += blockDim.x) src[e] = g[e];
    __syncthreads();
    resize_pass_h(src, tmp, H, W, R, kh, bh, ksh);
    __syncthreads();
    resize_pass_v(tmp, img + p.pad_t * P + colpad, P, R, kv, bv, ksv);
    __syncthreads();
    // first tower layer straight from the smem image: a warp per output row, a lane per CPL channels
    __nv_bfloat16* out = p.out + (b * 3 + br) * static_cast<int64_t>(p.Ho) * p.Wo * p.Cout + lane * CPL;
    for (int oy = warp; oy < p.Ho; oy += nw) {
      const float* row0 = img + (2 * oy) * P + coff;      // tap (kh = 0, ix = -pad_l) of pixel ox = 0
      __nv_bfloat16* orow = out + static_cast<int64_t>(oy) * p.Wo * p.Cout;
      for (int ox0 = 0; ox0 < p.Wo; ox0 += 4) {
        float x[3][9];
        if (coff == 0) {
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float* q = row0 + r * P + 2 * ox0;      // 32-byte aligned
            const float4 a = *reinterpret_cast<const float4*>(q);
            const float4 c4 = *reinterpret_cast<const float4*>(q + 4);
            x[r][0] = a.x; x[r][1] = a.y; x[r][2] = a.z; x[r][3] = a.w;
            x[r][4] = c4.x; x[r][5] = c4.y; x[r][6] = c4.z; x[r][7] = c4.w;
            x[r][8] = q[8];
          }
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cx = 0; cx < 9; ++cx) x[r][cx] = row0[r * P + 2 * ox0 + cx];
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          if (ox0 + px < p.Wo) {
            float acc[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) acc[c] = bs[c];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc[c] = fmaf(x[r][2 * px + kw], w[r * 3 + kw][c], acc[c]);
            uint32_t pk[CPL / 2];
#pragma unroll
            for (int c = 0; c < CPL; c += 2) {
              const __nv_bfloat162 h2 =
                  __floats2bfloat162_rn(apply_act(acc[c], p.act, p.alpha), apply_act(acc[c + 1], p.act, p.alpha));
              pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            __nv_bfloat16* o = orow + static_cast<int64_t>(ox0 + px) * p.Cout;
            if (CPL == 4) *reinterpret_cast<uint2*>(o) = make_uint2(pk[0], pk[CPL / 2 - 1]);
            else *reinterpret_cast<uint32_t*>(o) = pk[0];
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ K4 implicit GEMM
// Conv2D(3x3, strides 2, 'same') for Cin % 64 == 0 as a tcgen05 implicit GEMM:
//   M = 128 output pixels (a TH x Wo block of one image), N = Cout, K = 9 taps x Cin.
// The im2col operand is never materialised: for tap (kh,kw) and channel block cb the A tile is
// ONE 4-D tiled TMA load of the NHWC bf16 activation with element strides (1,2,2,1) — box
// {64 ch, 2*Wo, 2*TH, 1} starting at (cb*64, kw - pad, 2*oy0 + kh - pad, img) — which lands as
// TH*Wo rows of 128 B in the 128B-swizzled K-major layout UMMA reads; out-of-image taps are
// zero-filled by the TMA unit (TF 'same': the extra padding is after).  Weights [3*Cout][9*Cin]
// bf16 come through a 2-D map.  Same warp roles as k5_dense_stack; the epilogue adds the bias,
// applies ReLU / LeakyReLU and writes NHWC bf16.
constexpr int kCgThreads = 192;
constexpr int kCgMaxStages = 12;
// A ring stage holds the THREE taps of one kernel row (kw = 0,1,2) of one 64-channel block: three A
// tiles + three weight tiles, 12 UMMAs, ONE full/empty barrier round trip.  Measured (profiles/
// r1c_nets_experiments.txt): with one tap per stage the barrier protocol alone — no TMA, no MMA —
// cost 650 cycles per stage, 70 % of the kernel; the tensor floor of a stage is 64-128 cycles.
constexpr int kCgTapsPerStage = 3;
// share_kh (Wo % 8 == 0, Wo <= 32): taps kh = 0 and kh = 2 of a kernel column read the SAME input rows one
// output row apart (input rows 2 oy and 2 (oy + 1)), so a stage is (channel block, kw) and carries ONE box of
// TH + 1 even rows — kh = 2 is the same shared-memory tile entered Wo pixel rows (a multiple of the 1 024-byte
// swizzle period) further down — plus the TH odd rows of kh = 1 and the three weight tiles: 36 KB of activation
// per stage instead of 48 KB.  The kernel is bound by what the SMs pull out of L2 (432 KB per 128-pixel tile of
// sgan's 128 -> 64 layer, ~7.7 KB/clk over the chip), so this is time: 432 -> 360 KB per tile.
constexpr int kCgAEvenBytes = 160 * 128;     // share_kh: (TH + 1) * Wo <= 160 pixel rows of 128 B
struct ConvGemmParams {
  int stages;             // TMA ring depth (as many stages as fit)
  int share_kh;           // see above
  int64_t n_img;
  int Ho, Wo, Cin, Cout;
  int TH;                 // output rows per tile (TH * Wo <= 128)
  int tiles_per_img;
  int pad_t, pad_l;
  int act;
  float alpha;
  const float* bias;      // [3][Cout]
  __nv_bfloat16* out;     // [n_img][Ho][Wo][Cout]
};
__host__ __device__ constexpr int cg_stage_bytes(int cout, int share_kh = 0) {
  return share_kh ? kCgAEvenBytes + 128 * 128 + 3 * cout * 128 : kCgTapsPerStage * (128 * 128 + cout * 128);
}
__host__ __device__ constexpr int cg_pick_stages(int cout, int share_kh = 0) {
  return (232448 - 1024 - 256 - 3 * 128 * 4) / cg_stage_bytes(cout, share_kh) < kCgMaxStages
             ? (232448 - 1024 - 256 - 3 * 128 * 4) / cg_stage_bytes(cout, share_kh)
             : kCgMaxStages;
}
__host__ __device__ constexpr int cg_smem_bytes(int cout, int stages, int share_kh = 0) {
  return stages * cg_stage_bytes(cout, share_kh) + 1024 + 256 + 3 * 128 * 4;
}

__global__ void __launch_bounds__(kCgThreads, 1)
k4_conv_igemm(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const __grid_constant__ CUtensorMap map_xe, const ConvGemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int stage_bytes = cg_stage_bytes(p.Cout, p.share_kh);
  const int n_stages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + n_stages;
  uint64_t* tfull = empty + n_stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);     // (2 * 12 + 4) * 8 + 4 <= 256
  float* s_bias = reinterpret_cast<float*>(smem + n_stages * stage_bytes + 256);   // [3][Cout<=128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = p.n_img * p.tiles_per_img;
  const int cblocks = p.Cin / 64;
  const int k_stages = 3 * cblocks;                 // ring stages per tile: (kernel row, channel block)
  const int a_bytes = p.TH * p.Wo * 128;
  const int w_tile = p.Cout * 128;

  for (int e = threadIdx.x; e < 3 * p.Cout; e += blockDim.x) s_bias[e] = p.bias[e];
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    if (p.share_kh) tma_prefetch_desc(&map_xe);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // whole warp in uniform control flow, one elected lane issues (see elect_one())
    const uint64_t pol_a = policy_evict_last();    // every activation pixel is read ~2.25 times
    const uint64_t pol_b = policy_evict_last();
    uint32_t kit = 0, rs = 0, rph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t img = tile / p.tiles_per_img;
      const int tb = static_cast<int>(tile - img * p.tiles_per_img);
      int oy0 = tb * p.TH;
      if (oy0 > p.Ho - p.TH) oy0 = p.Ho - p.TH;        // last block overlaps instead of being ragged
      const int br = static_cast<int>(img % 3);
      if (p.share_kh) {
        const int ae_bytes = (p.TH + 1) * p.Wo * 128;
        for (int cb = 0; cb < cblocks; ++cb) {
          for (int kw = 0; kw < 3; ++kw, ++kit) {
            const int s = static_cast<int>(rs);           // ring slot and its phase, carried (no runtime division)
            mbar_wait(&empty[s], rph ^ 1u);
            if (++rs == static_cast<uint32_t>(n_stages)) { rs = 0; rph ^= 1u; }
            if (elect_one()) {
              unsigned char* a_dst = smem + s * stage_bytes;
              unsigned char* b_dst = a_dst + kCgAEvenBytes + 128 * 128;
              mbar_arrive_expect_tx(&full[s], ae_bytes + a_bytes + 3 * w_tile);
              // even input rows 2 oy0 .. 2 (oy0 + TH) (kh = 0 and, one output row down, kh = 2), odd rows (kh = 1)
              tma_load_4d(a_dst, &map_xe, cb * 64, kw - p.pad_l, 2 * oy0 - p.pad_t, static_cast<int32_t>(img),
                          &full[s], pol_a);
              tma_load_4d(a_dst + kCgAEvenBytes, &map_x, cb * 64, kw - p.pad_l, 2 * oy0 + 1 - p.pad_t,
                          static_cast<int32_t>(img), &full[s], pol_a);
#pragma unroll
              for (int kh = 0; kh < 3; ++kh)
                tma_load_2d(b_dst + kh * w_tile, &map_w, ((kh * 3 + kw) * cblocks + cb) * 64, br * p.Cout,
                            &full[s], pol_b);
            }
            __syncwarp();
          }
        }
        continue;
      }
      for (int kh = 0; kh < 3; ++kh) {
        for (int cb = 0; cb < cblocks; ++cb, ++kit) {
          const int s = static_cast<int>(rs);           // ring slot and its phase, carried (no runtime division)
          mbar_wait(&empty[s], rph ^ 1u);
          if (++rs == static_cast<uint32_t>(n_stages)) { rs = 0; rph ^= 1u; }
          if (elect_one()) {
            unsigned char* a_dst = smem + s * stage_bytes;
            unsigned char* b_dst = a_dst + kCgTapsPerStage * 128 * 128;
            mbar_arrive_expect_tx(&full[s], kCgTapsPerStage * (a_bytes + w_tile));
#pragma unroll
            for (int kw = 0; kw < kCgTapsPerStage; ++kw) {
              tma_load_4d(a_dst + kw * 128 * 128, &map_x, cb * 64, kw - p.pad_l, 2 * oy0 + kh - p.pad_t,
                          static_cast<int32_t>(img), &full[s], pol_a);
              tma_load_2d(b_dst + kw * w_tile, &map_w, ((kh * 3 + kw) * cblocks + cb) * 64, br * p.Cout,
                          &full[s], pol_b);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, p.Cout);
    uint32_t kit = 0, ait = 0, rs = 0, rph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ait) {
      const int ab = ait & 1;
      mbar_wait(&tempty[ab], ((ait >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * 128;
      for (int kb = 0; kb < k_stages; ++kb, ++kit) {
        const int s = static_cast<int>(rs);
        mbar_wait(&full[s], rph);
        if (++rs == static_cast<uint32_t>(n_stages)) { rs = 0; rph ^= 1u; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          if (p.share_kh) {
            const uint32_t b_addr = a_addr + kCgAEvenBytes + 128 * 128;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint32_t a_off = kh == 1 ? kCgAEvenBytes : (kh == 2 ? p.Wo * 128 : 0);
              const uint64_t da = umma_desc_k_sw128(a_addr + a_off);
              const uint64_t db = umma_desc_k_sw128(b_addr + kh * w_tile);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_f16(d_tmem, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc, (kb | kh | ks) != 0);
            }
            umma_commit(&empty[s]);
            if (kb == k_stages - 1) umma_commit(&tfull[ab]);
          } else {
          const uint32_t b_addr = a_addr + kCgTapsPerStage * 128 * 128;
#pragma unroll
          for (int kw = 0; kw < kCgTapsPerStage; ++kw) {
            const uint64_t da = umma_desc_k_sw128(a_addr + kw * 128 * 128);
            const uint64_t db = umma_desc_k_sw128(b_addr + kw * w_tile);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_f16(d_tmem, da + (ks * 32 >> 4), db + (ks * 32 >> 4), idesc, (kb | kw | ks) != 0);
          }
          umma_commit(&empty[s]);
          if (kb == k_stages - 1) umma_commit(&tfull[ab]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m / p.Wo, cx = m - r * p.Wo;
    const bool valid = r < p.TH;
    uint32_t ait = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ait) {
      const int ab = ait & 1;
      const int64_t img = tile / p.tiles_per_img;
      const int tb = static_cast<int>(tile - img * p.tiles_per_img);
      int oy0 = tb * p.TH;
      if (oy0 > p.Ho - p.TH) oy0 = p.Ho - p.TH;
      const float* bias = s_bias + (img % 3) * p.Cout;
      mbar_wait(&tfull[ab], (ait >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * 128 + (static_cast<uint32_t>(q * 32) << 16);
      __nv_bfloat16* out = p.out + ((img * p.Ho + oy0 + r) * p.Wo + cx) * p.Cout;
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float a0 = apply_act(__uint_as_float(v[e]) + bias[c0 + e], p.act, p.alpha);
            const float a1 = apply_act(__uint_as_float(v[e + 1]) + bias[c0 + e + 1], p.act, p.alpha);
            __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(out + c0);
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------ K5 dense stack
constexpr int kK5Threads = 192;
constexpr int kK5BlockM = 128;
constexpr int kK5N = 64;
constexpr int kK5BlockK = 64;        // bf16 elements = one 128-byte swizzle span
constexpr int kK5Stages = 8;
constexpr int kK5StageBytes = (kK5BlockM + kK5N) * kK5BlockK * 2;   // 24 576

struct K5Params {
  int64_t B;
  int kpg;                // 64-wide K blocks per ring stage (4, 2 or 1, dividing k_blocks): one barrier
                          // round trip (~650 cycles, see k4_conv_igemm) then covers 4 x 4 UMMAs
  int k_blocks;           // K / 64
  int C;                  // classes
  int head;               // 0 softmax (dnn, sgan c_model), 1 Z/(Z+1) (sgan d_model)
  int act1, act2;
  float alpha;
  const float* b1;        // [64]  (BN folded)
  const float* w2;        // [64][64] in-major
  const float* b2;
  const float* w3;        // [64][C]
  const float* b3;
  float* proba;           // [B][C]
  float* logits;          // [B][C], nullable
  int32_t* label;         // [B]
};

__host__ __device__ constexpr int k5_smem_bytes() {
  return kK5Stages * kK5StageBytes + 1024 + 256 + (64 * 64 + 64 * 8 + 64 * 2 + 8) * 4;
}

__global__ void __launch_bounds__(kK5Threads, 1)
k5_dense_stack(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_w1,
               const K5Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kK5Stages * kK5StageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kK5Stages;
  uint64_t* tfull = empty + kK5Stages;     // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_w2 = reinterpret_cast<float*>(smem + kK5Stages * kK5StageBytes + 256);
  float* s_w3 = s_w2 + 64 * 64;            // [64][8]
  float* s_b1 = s_w3 + 64 * 8;
  float* s_b2 = s_b1 + 64;
  float* s_b3 = s_b2 + 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.B + kK5BlockM - 1) / kK5BlockM;

  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) s_w2[e] = p.w2[e];
  for (int e = threadIdx.x; e < 64 * 8; e += blockDim.x) {
    const int i = e >> 3, c = e & 7;
    s_w3[e] = c < p.C ? p.w3[i * p.C + c] : 0.f;
  }
  if (threadIdx.x < 64) {
    s_b1[threadIdx.x] = p.b1[threadIdx.x];
    s_b2[threadIdx.x] = p.b2[threadIdx.x];
  }
  if (threadIdx.x < 8) s_b3[threadIdx.x] = threadIdx.x < p.C ? p.b3[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kK5Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&map_act);
    tma_prefetch_desc(&map_w1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ring: n_st stages of kpg K blocks each (the same bytes in flight for every kpg)
  const int kpg = p.kpg;
  const int n_st = kK5Stages / kpg;
  const int st_bytes = kpg * kK5StageBytes;
  const int k_steps = p.k_blocks / kpg;
  constexpr int kABytes = kK5BlockM * kK5BlockK * 2, kBBytes = kK5N * kK5BlockK * 2;
  if (warp == 0) {
    // whole warp in uniform control flow, one elected lane issues (see elect_one())
    const uint64_t pol_a = policy_evict_first();
    const uint64_t pol_b = policy_evict_last();
    uint32_t kit = 0, rs = 0, rph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int ks = 0; ks < k_steps; ++ks, ++kit) {
        const int s = static_cast<int>(rs);           // ring slot and its phase, carried (no runtime division)
        mbar_wait(&empty[s], rph ^ 1u);
        if (++rs == static_cast<uint32_t>(n_st)) { rs = 0; rph ^= 1u; }
        if (elect_one()) {
          unsigned char* a_dst = smem + s * st_bytes;
          unsigned char* b_dst = a_dst + kpg * kABytes;
          mbar_arrive_expect_tx(&full[s], st_bytes);
          for (int g = 0; g < kpg; ++g) {
            const int kb = ks * kpg + g;
            tma_load_2d(a_dst + g * kABytes, &map_act, kb * kK5BlockK, static_cast<int32_t>(tile * kK5BlockM),
                        &full[s], pol_a);
            tma_load_2d(b_dst + g * kBBytes, &map_w1, kb * kK5BlockK, 0, &full[s], pol_b);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc(kCF32, kFmtBF16, kFmtBF16, kK5BlockM, kK5N);
    uint32_t kit = 0, ait = 0, rs = 0, rph = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ait) {
      const int ab = ait & 1;
      mbar_wait(&tempty[ab], ((ait >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * kK5N;
      for (int ks = 0; ks < k_steps; ++ks, ++kit) {
        const int s = static_cast<int>(rs);
        mbar_wait(&full[s], rph);
        if (++rs == static_cast<uint32_t>(n_st)) { rs = 0; rph ^= 1u; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * st_bytes);
          const uint32_t b_addr = a_addr + kpg * kABytes;
          for (int g = 0; g < kpg; ++g) {
            const uint64_t da = umma_desc_k_sw128(a_addr + g * kABytes);
            const uint64_t db = umma_desc_k_sw128(b_addr + g * kBBytes);
#pragma unroll
            for (int k16 = 0; k16 < kK5BlockK / 16; ++k16)   // UMMA_K = 16 bf16 = 32 bytes
              umma_f16(d_tmem, da + (k16 * 32 >> 4), db + (k16 * 32 >> 4), idesc, (ks | g | k16) != 0);
          }
          umma_commit(&empty[s]);
          if (ks == k_steps - 1) umma_commit(&tfull[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t ait = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ait) {
      const int ab = ait & 1;
      const int64_t b = tile * kK5BlockM + m;
      mbar_wait(&tfull[ab], (ait >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * kK5N + (static_cast<uint32_t>(q * 32) << 16);
      float h1[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          h1[c0 + e] = apply_act(__uint_as_float(v[e]) + s_b1[c0 + e], p.act1, p.alpha);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
      if (b < p.B) {
        float h2[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) h2[j] = s_b2[j];
#pragma unroll 4
        for (int i = 0; i < 64; ++i) {
          const float x = h1[i];
          const float4* w4 = reinterpret_cast<const float4*>(s_w2 + i * 64);
#pragma unroll
          for (int j4 = 0; j4 < 16; ++j4) {
            const float4 wv = w4[j4];
            h2[j4 * 4 + 0] = fmaf(x, wv.x, h2[j4 * 4 + 0]);
            h2[j4 * 4 + 1] = fmaf(x, wv.y, h2[j4 * 4 + 1]);
            h2[j4 * 4 + 2] = fmaf(x, wv.z, h2[j4 * 4 + 2]);
            h2[j4 * 4 + 3] = fmaf(x, wv.w, h2[j4 * 4 + 3]);
          }
        }
        float lg[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) lg[c] = s_b3[c];
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float x = apply_act(h2[i], p.act2, p.alpha);
#pragma unroll
          for (int c = 0; c < 8; ++c) lg[c] = fmaf(x, s_w3[i * 8 + c], lg[c]);
        }
        int best = 0;
        float mx = lg[0];
#pragma unroll
        for (int c = 1; c < 8; ++c)
          if (c < p.C && lg[c] > mx) { mx = lg[c]; best = c; }
        if (p.head == 0) {
          double e[8], den = 0.0;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            e[c] = c < p.C ? exp(static_cast<double>(lg[c]) - static_cast<double>(mx)) : 0.0;
            den += e[c];
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < p.C) p.proba[b * p.C + c] = static_cast<float>(e[c] / den);
          p.label[b] = best;
        } else {
          double z = 0.0;   // sgan.py:125-129 custom_activation
#pragma unroll
          for (int c = 0; c < 8; ++c) z += c < p.C ? exp(static_cast<double>(lg[c])) : 0.0;
          const double d = z / (z + 1.0);
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < p.C) p.proba[b * p.C + c] = c == 0 ? static_cast<float>(d) : 0.f;
          p.label[b] = d >= 0.5 ? 1 : 0;
        }
        if (p.logits) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < p.C) p.logits[b * p.C + c] = lg[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace rml
