// Callers either side of the hot path:
//   k0_derive_targets  common.py:45-80 DerivedTarget.get_derived_targets — axis sums of the raw
//                      cube + top-k indices per axis (the SDK-free replacement for
//                      GetSensorTargets that yields the (i,j,k) SLICE mode needs)
//   k0_zoom_concat     common.py:143-149 with proj_zoom != 1: scipy.ndimage.zoom(order=3) is a
//                      fixed separable linear operator for given sizes; the host extracts the
//                      two operator matrices per projection from scipy and this kernel applies
//                      A_r . P . A_c^T in fp64, then concat + /255 like the identity path.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "k1_project.cuh"
#include "ptx.cuh"

namespace rml {

constexpr int kMaxTargets = 8;

struct DeriveParams {
  const float* cubes;
  int32_t* ijk;       // [B][T][3], ascending by axis sum like np.argsort (last = strongest)
  float* sums;        // nullable: [B][sx+sy+sz] axis sums (theta | phi | r)
  unsigned int* status;  // [3] += scans whose axis sums had no finite maximum
  int64_t B;
  int sx, sy, sz, T;
};

// one CTA (256 threads) per scan; a warp sums one (i,j) row at a time
__global__ void __launch_bounds__(256) k0_derive_targets(const DeriveParams p) {
  extern __shared__ float ds[];   // Sx[sx] | Sy[sy] | Sz[sz] | per-warp Sz partials [8][sz] | row sums
  float* Sx = ds;
  float* Sy = Sx + p.sx;
  float* Sz = Sy + p.sy;
  float* part = Sz + p.sz;
  float* rowsum = part + (blockDim.x >> 5) * p.sz;   // [sx*sy]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int rows = p.sx * p.sy;
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    for (int e = threadIdx.x; e < p.sx + p.sy + p.sz + nw * p.sz; e += blockDim.x) ds[e] = 0.f;
    __syncthreads();
    const float* cube = p.cubes + b * static_cast<int64_t>(rows) * p.sz;
    // four rows per warp step, all loads issued before the adds (memory-level parallelism)
    for (int r0 = warp * 4; r0 < rows; r0 += nw * 4) {
      float v[4][6];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* row = cube + static_cast<int64_t>(r0 + u) * p.sz;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int k = lane + 32 * q;
          v[u][q] = (r0 + u < rows && k < p.sz) ? row[k] : 0.f;
        }
      }
      if (p.sz <= 192) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            s += v[u][q];
            const int k = lane + 32 * q;
            if (k < p.sz) part[warp * p.sz + k] += v[u][q];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0 && r0 + u < rows) rowsum[r0 + u] = s;
        }
      } else {
        // wide rows: the generic strided walk
        for (int u = 0; u < 4 && r0 + u < rows; ++u) {
          const float* row = cube + static_cast<int64_t>(r0 + u) * p.sz;
          float s = 0.f;
          for (int k = lane; k < p.sz; k += 32) {
            const float x = row[k];
            s += x;
            part[warp * p.sz + k] += x;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) rowsum[r0 + u] = s;
        }
      }
    }
    __syncthreads();
    // fixed summation order: results are reproducible run to run
    for (int e = threadIdx.x; e < p.sx + p.sy + p.sz; e += blockDim.x) {
      float s = 0.f;
      if (e < p.sx) {
        for (int j = 0; j < p.sy; ++j) s += rowsum[e * p.sy + j];
        Sx[e] = s;
      } else if (e < p.sx + p.sy) {
        const int j = e - p.sx;
        for (int i = 0; i < p.sx; ++i) s += rowsum[i * p.sy + j];
        Sy[j] = s;
      } else {
        const int k = e - p.sx - p.sy;
        for (int w = 0; w < nw; ++w) s += part[w * p.sz + k];
        Sz[k] = s;
      }
    }
    __syncthreads();
    if (p.sums) {
      for (int e = threadIdx.x; e < p.sx + p.sy + p.sz; e += blockDim.x)
        p.sums[b * (p.sx + p.sy + p.sz) + e] = ds[e];
      __syncthreads();      // the ranking below overwrites the winners with -FLT_MAX
    }
    // top-T per axis: warp a (0..2) repeatedly takes the arg-max of its axis
    if (warp < 3) {
      float* S = warp == 0 ? Sx : (warp == 1 ? Sy : Sz);
      const int n = warp == 0 ? p.sx : (warp == 1 ? p.sy : p.sz);
      for (int t = 0; t < p.T; ++t) {
        float best = -FLT_MAX;
        int bi = 0x7fffffff;
        for (int e = lane; e < n; e += 32) {
          const float v = S[e];
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
          if (bi >= n) {
            // no candidate compared greater than -FLT_MAX (NaN or -inf sums): take the lowest
            // unused index instead of the sentinel and report it (status[3] -> RML_E_INVALID)
            bi = 0;
            while (bi < n - 1 && S[bi] == -FLT_MAX) ++bi;
            if (p.status) atomicAdd(p.status + 3, 1u);
          }
          // rank t from the top goes to slot T-1-t (ascending order of the reference)
          p.ijk[(b * p.T + (p.T - 1 - t)) * 3 + warp] = bi;
          S[bi] = -FLT_MAX;
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ zoom
struct ZoomParams {
  const float* proj[3];   // per projection: first scan's [ih][iw] block (nullable when masked)
  int64_t pstride[3];     // elements between consecutive scans of that projection
  int ih[3], iw[3], oh[3], ow[3];
  const double* ar[3];    // [oh][ih]
  const double* ac[3];    // [ow][iw]
  int off[3];             // output offsets inside the feature row
  float* feats;           // [B][F]
  int64_t B;
  int F;
  int scale;
  float offset, scale_value;
  const float* aff_off;   // nullable per-feature tables (indexed by output feature)
  const float* aff_scl;
};

// grid (scans, 3); smem: P [ih][iw] f32 | T [ih][ow] f64
__global__ void __launch_bounds__(256) k0_zoom_concat(const ZoomParams p) {
  extern __shared__ double zs[];
  const int q = blockIdx.y;
  if (!p.proj[q]) return;
  const int ih = p.ih[q], iw = p.iw[q], oh = p.oh[q], ow = p.ow[q];
  double* T = zs;                                            // [ih][ow]
  float* P = reinterpret_cast<float*>(zs + ih * ow);         // [ih][iw]
  const double* ar = p.ar[q];
  const double* ac = p.ac[q];
  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* src = p.proj[q] + b * p.pstride[q];
    for (int e = threadIdx.x; e < ih * iw; e += blockDim.x) P[e] = src[e];
    __syncthreads();
    for (int e = threadIdx.x; e < ih * ow; e += blockDim.x) {
      const int r = e / ow, oc = e - r * ow;
      double s = 0.0;
      for (int c = 0; c < iw; ++c) s = fma(static_cast<double>(P[r * iw + c]), ac[oc * iw + c], s);
      T[e] = s;
    }
    __syncthreads();
    float* out = p.feats + b * p.F + p.off[q];
    for (int e = threadIdx.x; e < oh * ow; e += blockDim.x) {
      const int orow = e / ow, oc = e - orow * ow;
      double s = 0.0;
      for (int r = 0; r < ih; ++r) s = fma(ar[orow * ih + r], T[r * ow + oc], s);
      const float v = static_cast<float>(s);                 // ndimage.zoom returns the input dtype
      out[e] = p.scale ? affine_apply(v, p.offset, p.scale_value, p.aff_off, p.aff_scl, p.off[q] + e) : v;
    }
    __syncthreads();
  }
}

}  // namespace rml
