// Convolution entry points (device pointers), see conv.cu.
#pragma once
#include "common.cuh"

namespace bdsp {

// largest L the overlap-save kernels support
template <typename T> size_t ols_max_taps();

// Overlap-save plan of one impulse response (the reference recomputes this spectrum inside every
// convolve_signal call, convolution.rs:333-339,419-425).  Immutable after ols_plan_create, so any number of
// threads / streams of the creating device may convolve with it concurrently.  Destruction synchronises the device.
struct OlsPlan {
    int is64 = 0, device = 0, h_is_real = 0;
    size_t L = 0;
    // primary block: M points; Hs = FFT_M(pad(h))/M in natural order (M complex) followed, for the fused c32 kernels
    // (M = 4096 / 8192), by the same spectrum in the kernel's layout (2*M floats), bound to a texture object
    size_t M = 0;
    void* Hs = nullptr;
    cudaTextureObject_t htex = 0;
    // c32, 2 <= L <= 2046: second fused block length (8192) next to M = 4096; chosen per call from the amount of work
    size_t M2 = 0;
    void* Hs2 = nullptr;
    cudaTextureObject_t htex2 = 0;
    // c32 responses longer than the fused blocks: the plan's own copy of the taps (L real or complex values) for the
    // full-length frequency-domain path, which power-of-two vectors take instead of the generic blocks
    void* taps_dev = nullptr;
};
// h: L complex (or real) taps on the device.  complex_signal: the vectors that will be convolved are complex.
template <typename T> OlsPlan* ols_plan_create(const void* h, size_t L, int h_is_real, bool complex_signal, cudaStream_t st);
void ols_plan_destroy(OlsPlan* p);
// y <- centred circular convolution of every one of `batch` vectors of N points (x != y)
template <typename T> int ols_plan_convolve(const OlsPlan* p, const void* x, void* y, size_t N, size_t batch, int is_real, cudaStream_t st);
// direct form, y[i] = sum_k x[(i + cl - 1 - k) mod N] h[k]
template <typename T>
int fir_convolve(const void* x, void* y, const void* h, size_t N, size_t batch, size_t L, size_t cl, int x_complex,
                 int h_complex, cudaStream_t st);
// full-length FFT path (any N, any L <= N)
template <typename T>
int fft_convolve_full(const void* x, void* y, const void* h, size_t N, size_t batch, size_t L, int is_real,
                      int h_is_real, cudaStream_t st);

}  // namespace bdsp
