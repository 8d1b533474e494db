// Host-side lossless narrowing of integral float32 cubes (plain C++, compiled by the host compiler).
//
// predict.py:90-91 widens the Walabot's integer voxels (0..255) with np.array(raw_image, dtype=np.float32).
// rml_predict_host is PCIe bound on those float32 cubes (53 of 55.5 GB/s on the B200 box), so for a
// model on the integer path it converts each chunk back to bytes on the host — a pool of threads, AVX2,
// every value checked to be exactly an integer in [0, 255] — and moves a quarter of the bytes over the
// bus; the device then runs the uint8-cube kernels, whose results are bit-identical to the float32 ones
// (tests/test_gpu_u8cubes.py).  A chunk with any other value (NaN included) is copied as float32, exactly
// as before.  Nothing here touches the GPU.  Measured on the B200 box (16 vCPUs): 65-90 GB/s of float32
// input, i.e. host-memory bound at about the rate PCIe 5 x16 moves the float32 bytes anyway (+3 % end to
// end), so rml_predict_host uses it only on request (rml_set_host_narrowing / RML_HOST_NARROW=1).
#pragma once
#include <cstddef>
#include <cstdint>

namespace rml_host {

struct NarrowPool;   // fork-join pool of conversion threads

// threads <= 0: one per CPU the process may run on (sched_getaffinity), at most 32
NarrowPool* narrow_pool_create(int threads);
void narrow_pool_destroy(NarrowPool* p);
int narrow_pool_threads(const NarrowPool* p);
// dst[i] = (uint8_t)src[i] for i < n; returns 0 when every src[i] is exactly an integer in [0, 255],
// 1 otherwise (dst is then unspecified).  dst must be 32-byte aligned.
int narrow_f32_to_u8(NarrowPool* p, const float* src, uint8_t* dst, size_t n);
// single-threaded reference of the same conversion (tests)
int narrow_f32_to_u8_scalar(const float* src, uint8_t* dst, size_t n);

}  // namespace rml_host
