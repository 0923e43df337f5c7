// Convolution entry points (device pointers), see conv.cu.
#pragma once
#include "common.cuh"

namespace bdsp {

// block length the overlap-save kernel uses for an L-tap impulse response
template <typename T> size_t ols_block_len(size_t L, bool complex_signal);
// largest L the overlap-save kernel supports
template <typename T> size_t ols_max_taps();
// bytes the caller must provide for Hs
template <typename T> size_t ols_spectrum_bytes(size_t M);
// Hs (ols_spectrum_bytes(M) bytes) <- FFT_M(pad(h)) / M (+ the permuted copy for the fused kernel).  h: L complex (or real) taps on the device.
template <typename T> int ols_prepare(const void* h, size_t L, int h_is_real, void* Hs, size_t M, cudaStream_t st);
// y <- centred circular convolution of every one of `batch` vectors of N points with h (via Hs)
template <typename T>
int ols_convolve(const void* x, void* y, size_t N, size_t batch, size_t L, const void* Hs, size_t M, int is_real,
                 cudaStream_t st);
// direct form, y[i] = sum_k x[(i + cl - 1 - k) mod N] h[k]
template <typename T>
int fir_convolve(const void* x, void* y, const void* h, size_t N, size_t batch, size_t L, size_t cl, int x_complex,
                 int h_complex, cudaStream_t st);
// full-length FFT path (any N, any L <= N)
template <typename T>
int fft_convolve_full(const void* x, void* y, const void* h, size_t N, size_t batch, size_t L, int is_real,
                      int h_is_real, cudaStream_t st);

}  // namespace bdsp
