// Host-side FFT entry point shared by the C ABI and the convolution code.
#pragma once
#include "common.cuh"

namespace bdsp {

struct FftOpts {
    int inverse = 0;        // 0: exp(-i...), 1: exp(+i...); never normalised (rustfft semantics)
    int real_input = 0;     // input holds n real scalars per sequence (imag = 0), output complex
    size_t in_rot = 0;      // read sequence element (i + in_rot) mod n   (ifft_shift fused on load)
    size_t out_rot = 0;     // write result k to (k + out_rot) mod n      (fft_shift fused on store)
    double scale = 1.0;     // multiply input by scale while loading (ifft's 1/points)
    int magnitude = 0;      // epilogue: store hypot(re, im) as n real scalars per sequence
    InMul in_mul;           // multiply input element g of every sequence by a table entry / window value while loading
                            // (windowed_fft: time_to_freq.rs:167-175; spectrum multiply of the frequency-domain
                            // convolution: convolution.rs:427-429,444-446) - no separate pass over memory
};

// Transforms `batch` sequences of n complex points (sequence b at element offset b*n; real input:
// b*n scalars).  `work` must hold n*batch complex values when the transform needs a scratch pass
// (n > block limit or non power of two); pass nullptr to use the library's grow-only workspace.
// in == out is allowed.  Returns 0 or a negative error code (see set_last_error()).
template <typename T>
int fft_exec(const void* in, void* out, size_t n, size_t batch, const FftOpts& opts, void* work,
             size_t work_bytes, cudaStream_t stream);

// largest n the single-CTA shared-memory kernel handles for this precision
template <typename T> size_t fft_block_max_n();

// device pointer to the master twiddle table W_{2^14}^i for the current device
template <typename T> const typename CpxOf<T>::type* twiddle_table();

// grow-only per-device workspace (device memory); not thread safe across host threads sharing a device
void* workspace(size_t bytes, int slot);
void workspace_bind_stream(cudaStream_t st);
void workspace_release_stream(cudaStream_t st);

}  // namespace bdsp
