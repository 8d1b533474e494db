// Stand-alone sm_100a probe (not part of the library): answers three questions the fused tower
// kernel depends on, on real hardware, in one gpurun call.
//   A. tcgen05.mma shared-memory descriptor, K-major SWIZZLE_128B, start address 128-byte aligned
//      but NOT 1024-byte aligned (rows r..r+127 of a longer array of 128-byte rows): which of
//      {absolute-address swizzle, start-relative swizzle} x {base_offset = 0, base_offset = r & 7}
//      multiplies correctly?
//   B. K-major SWIZZLE_NONE descriptor for a K = 32 (64-byte row) operand: which field is the
//      K-direction core-matrix stride and which the M-direction one?
//   C. cycle costs with clock64(): mbarrier.try_wait on a completed phase, arrive.expect_tx,
//      tcgen05.commit -> mbarrier visible, issue cost of a batch of UMMAs, UMMA batch completion,
//      tcgen05.ld throughput (4 warps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../radar_ml_b200/csrc/ptx.cuh"

using namespace rml;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

struct ProbeOut {
  float d[8][128 * 64];       // results of up to 8 variants, [m][n]
  long long t[32];
};

// ---- test A: variant v = 2*rel + use_bo, start row r
// smem A: 144 rows x 128 B (64 bf16 each); B: 32 rows x 128 B standard swizzle.
__global__ void __launch_bounds__(128, 1) probe_a(const __nv_bfloat16* A /*[144][64]*/, const __nv_bfloat16* Bm /*[32][64]*/,
                                                   ProbeOut* out, int r) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                  // 144 * 128 = 18432
  unsigned char* sB = smem + 18432;          // 32 * 128 = 4096
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 18432 + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  for (int v = 0; v < 4; ++v) {
    const int rel = v >> 1, use_bo = v & 1;
    // fill: logical row i, chunk c (16 B = 8 bf16)
    for (int e = threadIdx.x; e < 144 * 8; e += blockDim.x) {
      const int i = e >> 3, c = e & 7;
      const int sw = rel ? ((i - r) & 7) : (i & 7);
      *reinterpret_cast<uint4*>(sA + i * 128 + ((c ^ sw) << 4)) = *reinterpret_cast<const uint4*>(A + i * 64 + c * 8);
    }
    for (int e = threadIdx.x; e < 32 * 8; e += blockDim.x) {
      const int i = e >> 3, c = e & 7;
      *reinterpret_cast<uint4*>(sB + i * 128 + ((c ^ (i & 7)) << 4)) = *reinterpret_cast<const uint4*>(Bm + i * 64 + c * 8);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, 32);
      const uint32_t a0 = smem_u32(sA) + r * 128;
      const uint64_t da = desc_sw128(a0, use_bo ? (r & 7) : 0);
      const uint64_t db = desc_sw128(smem_u32(sB), 0);
      for (int k = 0; k < 4; ++k) umma_f16(tmem, da + (k * 32 >> 4), db + (k * 32 >> 4), idesc, k != 0);
      umma_commit(bar);
    }
    mbar_wait(bar, v & 1);
    tc_fence_after();
    uint32_t vals[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), vals);
    tmem_ld_wait();
    const int m = threadIdx.x;
    for (int n = 0; n < 32; ++n) out->d[v][m * 64 + n] = __uint_as_float(vals[n]);
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ---- test B: no-swizzle K-major, A [128][32] bf16, B [64][32] bf16; element (m,k) at
// (m/8)*512 + (k/8)*128 + (m%8)*16 + (k%8)*2.  variant 0: desc.lbo = 128 (K stride), sbo = 512; variant 1: swapped
__global__ void __launch_bounds__(128, 1) probe_b(const __nv_bfloat16* A /*[128][32]*/, const __nv_bfloat16* Bm /*[64][32]*/,
                                                   ProbeOut* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                  // 128 * 64 = 8192
  unsigned char* sB = smem + 8192;           // 64 * 64 = 4096
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  for (int e = threadIdx.x; e < 128 * 4; e += blockDim.x) {
    const int m = e >> 2, kc = e & 3;
    *reinterpret_cast<uint4*>(sA + (m >> 3) * 512 + kc * 128 + (m & 7) * 16) = *reinterpret_cast<const uint4*>(A + m * 32 + kc * 8);
  }
  for (int e = threadIdx.x; e < 64 * 4; e += blockDim.x) {
    const int m = e >> 2, kc = e & 3;
    *reinterpret_cast<uint4*>(sB + (m >> 3) * 512 + kc * 128 + (m & 7) * 16) = *reinterpret_cast<const uint4*>(Bm + m * 32 + kc * 8);
  }
  fence_proxy_async_smem();
  __syncthreads();
  for (int v = 0; v < 2; ++v) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, 64);
      for (int k = 0; k < 2; ++k) {     // K = 16 per instruction = 2 core matrices along K
        const uint64_t da = v == 0 ? desc_noswz(smem_u32(sA) + k * 256, 128, 512) : desc_noswz(smem_u32(sA) + k * 256, 512, 128);
        const uint64_t db = v == 0 ? desc_noswz(smem_u32(sB) + k * 256, 128, 512) : desc_noswz(smem_u32(sB) + k * 256, 512, 128);
        umma_f16(tmem, da, db, idesc, k != 0);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, v & 1);
    tc_fence_after();
    uint32_t vals[32];
    const int m = threadIdx.x;
    for (int h = 0; h < 2; ++h) {
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + h * 32, vals);
      tmem_ld_wait();
      for (int n = 0; n < 32; ++n) out->d[4 + v][m * 64 + h * 32 + n] = __uint_as_float(vals[n]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// ---- test C: timings (single CTA, 128 threads)
__global__ void __launch_bounds__(128, 1) probe_c(ProbeOut* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                  // 128 rows x 128 B
  unsigned char* sB = smem + 16384;          // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < (16384 + 8192) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    long long t0, t1;
    // 1. arrive + try_wait on the completed phase, 64 round trips on one barrier
    t0 = clock64();
    for (int i = 0; i < 64; ++i) { mbar_arrive(&bar[0]); mbar_wait(&bar[0], i & 1); }
    t1 = clock64();
    out->t[0] = (t1 - t0) / 64;
    // 2. tcgen05.commit with nothing pending -> wait, 64 round trips
    t0 = clock64();
    for (int i = 0; i < 64; ++i) { umma_commit(&bar[1]); mbar_wait(&bar[1], i & 1); }
    t1 = clock64();
    out->t[1] = (t1 - t0) / 64;
    // 3. issue cost of 36 UMMAs (M=128, N=32, K=16) and completion
    const uint32_t idesc32 = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, 32);
    const uint32_t idesc64 = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, 64);
    const uint64_t da = umma_desc_k_sw128(smem_u32(sA));
    const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int i = 0; i < 36; ++i) umma_f16(tmem, da + ((i & 3) * 32 >> 4), db + ((i & 3) * 32 >> 4), idesc32, i != 0);
      t1 = clock64();
      umma_commit(&bar[2]);
      mbar_wait(&bar[2], rep & 1);
      long long t2 = clock64();
      out->t[2 + rep * 2] = t1 - t0;       // issue
      out->t[3 + rep * 2] = t2 - t0;       // issue + completion
    }
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int i = 0; i < 36; ++i) umma_f16(tmem + 64, da + ((i & 3) * 32 >> 4), db + ((i & 3) * 32 >> 4), idesc64, i != 0);
      t1 = clock64();
      umma_commit(&bar[3]);
      mbar_wait(&bar[3], rep & 1);
      long long t2 = clock64();
      out->t[6 + rep * 2] = t1 - t0;
      out->t[7 + rep * 2] = t2 - t0;
    }
    // 4. 144 UMMAs N=64 (4 x 36): does completion scale with the 32-cycle floor?
    t0 = clock64();
    for (int i = 0; i < 144; ++i) umma_f16(tmem + 64, da + ((i & 3) * 32 >> 4), db + ((i & 3) * 32 >> 4), idesc64, i != 0);
    t1 = clock64();
    umma_commit(&bar[2]);
    mbar_wait(&bar[2], 0);
    out->t[10] = t1 - t0;
    out->t[11] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  // 5. tcgen05.ld throughput: every warp reads its 32 lanes x 256 columns, 8 times
  {
    uint32_t v[32];
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int rep = 0; rep < 8; ++rep)
      for (int c = 0; c < 256; c += 32) {
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        acc += v[0] ^ v[31];
      }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out->t[12] = t1 - t0;      // cycles for 8 x 128 lanes x 256 cols x 4 B = 1 MiB
    if (acc == 0x12345678) out->t[31] = acc;
    // same without the wait after every load (4 loads in flight)
    __syncthreads();
    const long long t2 = clock64();
    for (int rep = 0; rep < 8; ++rep)
      for (int c = 0; c < 256; c += 128) {
        uint32_t a[32], b[32], c2[32], d2[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, a);
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c + 32, b);
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c + 64, c2);
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c + 96, d2);
        tmem_ld_wait();
        acc += a[0] ^ b[31] ^ c2[5] ^ d2[7];
      }
    __syncthreads();
    const long long t3 = clock64();
    if (threadIdx.x == 0) out->t[13] = t3 - t2;
    if (acc == 0x12345678) out->t[31] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- test D: the same 36 UMMAs issued by an elected lane of a converged warp (uniform control flow)
__global__ void __launch_bounds__(128, 1) probe_d(ProbeOut* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < 32768 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384);
    for (int variant = 0; variant < 2; ++variant) {
      const uint32_t idesc = umma_idesc(kCF32, kFmtBF16, kFmtBF16, 128, variant ? 64 : 32);
      for (int it = 0; it < 3; ++it) {
        const long long t0 = clock64();
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t da = umma_desc_k_sw128(a0);
            const uint64_t db = umma_desc_k_sw128(b0 + (tap & 1) * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem, da + (k * 32 >> 4), db + (k * 32 >> 4), idesc, (tap | k) != 0);
          }
          umma_commit(&bar[0]);
        }
        __syncwarp();
        const long long t1 = clock64();
        mbar_wait(&bar[0], (variant * 3 + it) & 1);
        const long long t2 = clock64();
        if (threadIdx.x == 32) { out->t[16 + variant * 6 + it * 2] = t1 - t0; out->t[17 + variant * 6 + it * 2] = t2 - t0; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
  CK(cudaSetDevice(0));
  ProbeOut* out;
  CK(cudaMallocManaged(&out, sizeof(ProbeOut)));
  memset(out, 0, sizeof(ProbeOut));
  std::vector<__nv_bfloat16> A(144 * 64), B(32 * 64), A2(128 * 32), B2(64 * 32);
  std::vector<float> Af(144 * 64), Bf(32 * 64), A2f(128 * 32), B2f(64 * 32);
  srand(7);
  auto fill = [](std::vector<__nv_bfloat16>& h, std::vector<float>& f) {
    for (size_t i = 0; i < h.size(); ++i) { float v = static_cast<float>(rand() % 9 - 4); f[i] = bf(v); h[i] = __float2bfloat16(v); }
  };
  fill(A, Af); fill(B, Bf); fill(A2, A2f); fill(B2, B2f);
  __nv_bfloat16 *dA, *dB, *dA2, *dB2;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dA2, A2.size() * 2)); CK(cudaMalloc(&dB2, B2.size() * 2));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA2, A2.data(), A2.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB2, B2.data(), B2.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  CK(cudaFuncSetAttribute(probe_b, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  CK(cudaFuncSetAttribute(probe_c, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  const int rs[] = {0, 1, 3, 8, 13};
  for (int r : rs) {
    probe_a<<<1, 128, 40960>>>(dA, dB, out, r);
    CK(cudaDeviceSynchronize());
    for (int v = 0; v < 4; ++v) {
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += Af[(r + m) * 64 + k] * Bf[n * 64 + k];
          const double e = fabs(ref - out->d[v][m * 64 + n]);
          if (e > maxerr) maxerr = e;
        }
      printf("A: start row r=%2d  swizzle=%s  base_offset=%d  max|err|=%g  %s\n", r, (v >> 1) ? "start-relative" : "absolute-addr ",
             (v & 1) ? (r & 7) : 0, maxerr, maxerr == 0 ? "OK" : "WRONG");
    }
  }
  probe_b<<<1, 128, 40960>>>(dA2, dB2, out);
  CK(cudaDeviceSynchronize());
  for (int v = 0; v < 2; ++v) {
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        float ref = 0;
        for (int k = 0; k < 32; ++k) ref += A2f[m * 32 + k] * B2f[n * 32 + k];
        const double e = fabs(ref - out->d[4 + v][m * 64 + n]);
        if (e > maxerr) maxerr = e;
      }
    printf("B: no-swizzle K-major, desc (lbo,sbo) = %s  max|err|=%g  %s\n", v == 0 ? "(K stride 128, M stride 512)" : "(512, 128) swapped",
           maxerr, maxerr == 0 ? "OK" : "WRONG");
  }
  probe_c<<<1, 128, 40960>>>(out);
  CK(cudaDeviceSynchronize());
  printf("C: mbarrier arrive + try_wait round trip        %lld cycles\n", out->t[0]);
  printf("C: tcgen05.commit (idle) + try_wait round trip   %lld cycles\n", out->t[1]);
  printf("C: 36 UMMA M128 N32 K16: issue %lld / done %lld cycles (cold), issue %lld / done %lld (warm); floor 36*16 = 576\n",
         out->t[2], out->t[3], out->t[4], out->t[5]);
  printf("C: 36 UMMA M128 N64 K16: issue %lld / done %lld cycles (cold), issue %lld / done %lld (warm); floor 36*32 = 1152\n",
         out->t[6], out->t[7], out->t[8], out->t[9]);
  printf("C: 144 UMMA M128 N64 K16: issue %lld / done %lld cycles; floor 4608\n", out->t[10], out->t[11]);
  printf("C: tcgen05.ld 1 MiB by 4 warps, wait after each x32: %lld cycles (%.1f B/clk); 4 in flight: %lld cycles (%.1f B/clk)\n",
         out->t[12], 1048576.0 / out->t[12], out->t[13], 1048576.0 / out->t[13]);
  CK(cudaFuncSetAttribute(probe_d, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  probe_d<<<1, 128, 40960>>>(out);
  CK(cudaDeviceSynchronize());
  for (int v = 0; v < 2; ++v)
    printf("D: elect_one() issue, 36 UMMA M128 N%d K16: issue %lld / done %lld, %lld / %lld, %lld / %lld cycles (floor %d)\n", v ? 64 : 32,
           out->t[16 + v * 6], out->t[17 + v * 6], out->t[18 + v * 6], out->t[19 + v * 6], out->t[20 + v * 6], out->t[21 + v * 6], v ? 1152 : 576);
  return 0;
}
