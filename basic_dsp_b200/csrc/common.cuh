// Common helpers for the basic_dsp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define BDSP_SM_COUNT_DEFAULT 148

namespace bdsp {

// --- last error (thread local), reported through bdsp_last_error() ------------------------
void set_last_error(const char* fmt, ...);
const char* get_last_error();
void count_launch();
unsigned long long launch_count();

#define BDSP_CUDA_OK(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            bdsp::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                                 cudaGetErrorString(_e));                                   \
            return -1000 - (int)_e;                                                         \
        }                                                                                   \
    } while (0)

// one kernel launch was just issued: count it and surface launch-configuration errors
#define BDSP_LAUNCHED()                                                                     \
    do {                                                                                    \
        bdsp::count_launch();                                                               \
        BDSP_CUDA_OK(cudaGetLastError());                                                   \
    } while (0)

// workspace(): nullptr (and last error set) when the allocation fails
#define BDSP_WS(ptr, T_, bytes, slot)                                                       \
    T_ ptr = reinterpret_cast<T_>(bdsp::workspace((bytes), (slot)));                        \
    if (!ptr) return -1001

// --- complex scalar types ------------------------------------------------------------------
template <typename T> struct CpxOf;
template <> struct CpxOf<float> { typedef float2 type; };
template <> struct CpxOf<double> { typedef double2 type; };

template <typename T> __host__ __device__ __forceinline__ typename CpxOf<T>::type mk(T re, T im) {
    typename CpxOf<T>::type r;
    r.x = re;
    r.y = im;
    return r;
}

template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// complex product; FMA contraction allowed (transform kernels, tolerance is rel-L2)
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }

// exact-rounding product (ac-bd, ad+bc): every operation individually rounded, no FMA.  This is
// what the reference's CPU path computes (num-complex Mul, simd_extensions/fallback.rs:195-203).
__device__ __forceinline__ float2 cmul_nofma(float2 a, float2 b) {
    float2 r;
    r.x = __fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    r.y = __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
    return r;
}
__device__ __forceinline__ double2 cmul_nofma(double2 a, double2 b) {
    double2 r;
    r.x = __dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
    r.y = __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
    return r;
}

__device__ __forceinline__ void sincospi_t(float x, float* s, float* c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double* s, double* c) { sincospi(x, s, c); }

// exp(sign * 2*pi*i * num/den), argument reduced exactly in integers first
template <typename T>
__device__ __forceinline__ typename CpxOf<T>::type unit_root(unsigned long long num, unsigned long long den, int sign) {
    num %= den;
    T s, c;
    sincospi_t((T)(2.0 * (double)num / (double)den), &s, &c);
    return mk<T>(c, sign < 0 ? -s : s);
}

static inline int ilog2(size_t v) {
    int l = 0;
    while ((1ull << (l + 1)) <= v) l++;
    return l;
}
static inline bool is_pow2(size_t v) { return v && !(v & (v - 1)); }
static inline size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

int sm_count();

}  // namespace bdsp
