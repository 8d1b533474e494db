// B200 probe: latency and throughput of DFMA and F2F.F64.F32 (the resize role of k6_tower is built on them)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_lat(double* out, long long* cyc, int n) {
  double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = a * b + c;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void dfma_thr(double* out, long long* cyc, int n) {
  double a[CH];
  for (int k = 0; k < CH; ++k) a[k] = threadIdx.x * 1e-3 + k;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k) a[k] = a[k] * b + c;
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < CH; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void cvt_thr(double* out, long long* cyc, int n, const float* in) {
  double a[CH];
  float f[CH];
  for (int k = 0; k < CH; ++k) { a[k] = 0; f[k] = in[threadIdx.x + k]; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < CH; ++k) { a[k] += static_cast<double>(f[k]); f[k] = __int_as_float(__float_as_int(f[k]) ^ (i & 1)); }
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < CH; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc; float* in;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096); cudaMalloc(&in, 1 << 16);
  cudaMemset(in, 0, 1 << 16);
  long long h;
  const int n = 4096;
  dfma_lat<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent chain, 1 warp: %.2f cycles per DFMA (latency)\n", double(h) / n);
  for (int warps : {1, 4, 8, 16, 32}) {
    dfma_thr<8><<<1, 32 * warps>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 chains/thread, %2d warps on one SM: %.2f cycles per warp-DFMA per SM -> %.1f lanes/clk/SM\n", warps,
           double(h) / (n * 8.0 * warps), 32.0 * n * 8 * warps / double(h));
  }
  for (int warps : {1, 4, 8, 16}) {
    cvt_thr<8><<<1, 32 * warps>>>(out, cyc, n, in); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("F2F.F64.F32 + DADD + LOP, 8 chains/thread, %2d warps: %.2f cycles per warp-(cvt,add) per SM\n", warps,
           double(h) / (n * 8.0 * warps));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
