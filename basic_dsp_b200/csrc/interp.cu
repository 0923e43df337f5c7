// Interpolation kernels: interpolatef (polyphase FIR, integer factor; per-output taps otherwise)
// and interpolate_lin.  Reference: vector/src/vector_types/time_freq/interpolation.rs:92-482 and
// real_interpolation.rs:33-71.
#include "interp.cuh"

namespace bdsp {

// ------------------------------------------------------------------------------------------
// Integer-factor polyphase kernel.
//   y[r*F + s] = sum_{j < J} x[(r - L - 1 + j) mod N] * tab[sel][s][j],   J = 2L + 3
// `tab` holds two tables over the same window superset [r-L-1, r+L+1]:
//   sel 0 = interior outputs  (interpolate_priv_simd, interpolation.rs:244-275; taps reversed)
//   sel 1 = edge outputs      (interpolate_priv_simd_step, :293-315)
// built on the host in precision T exactly as function_to_vectors (:133-181) evaluates them.
// One thread = one input position r = all F phases; the x window is read once per thread from a
// shared-memory tile (consecutive threads -> consecutive addresses), the taps are warp-uniform
// broadcast loads.
// ------------------------------------------------------------------------------------------
#define IP_THREADS 256

template <typename T, bool CPLX, int FMAX>
__global__ void __launch_bounds__(IP_THREADS)
interp_poly_kernel(const void* __restrict__ x_, void* __restrict__ y_, const T* __restrict__ tab, long long N,
                   long long new_points, int F, int L, long long scalar_len) {
    typedef typename CpxOf<T>::type C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int J = 2 * L + 3;
    T* stab = reinterpret_cast<T*>(smem_raw);                 // [2][F][J]
    const int tab_elems = 2 * F * J;
    const int tab_pad = (tab_elems + 3) & ~3;
    typedef typename std::conditional<CPLX, C, T>::type X;
    X* sx = reinterpret_cast<X*>(stab + tab_pad);             // IP_THREADS + J - 1 window
    const long long r0 = (long long)blockIdx.x * IP_THREADS;
    for (int i = threadIdx.x; i < tab_elems; i += IP_THREADS) stab[i] = tab[i];
    const int W = IP_THREADS + J - 1;
    long long g = (r0 - L - 1 + threadIdx.x) % N;
    if (g < 0) g += N;
    const long long adv = IP_THREADS % N;
    for (int w = threadIdx.x; w < W; w += IP_THREADS) {
        sx[w] = reinterpret_cast<const X*>(x_)[g];
        g += adv; if (g >= N) g -= N;
    }
    __syncthreads();
    const long long r = r0 + threadIdx.x;
    if (r * F >= new_points) return;
    X acc[FMAX];
    int sel[FMAX];
#pragma unroll
    for (int s = 0; s < FMAX; s++) {
        long long i = r * F + s;
        bool interior = (i >= scalar_len) && (i < new_points - scalar_len);
        sel[s] = (interior ? 0 : F * J) + s * J;
        if (CPLX) { acc[s] = X(); }
        acc[s] = X();
    }
    for (int j = 0; j < J; j++) {
        const X xv = sx[threadIdx.x + j];
#pragma unroll
        for (int s = 0; s < FMAX; s++) {
            if (s < F) {
                const T t = stab[sel[s] + j];
                if constexpr (CPLX) { acc[s].x += xv.x * t; acc[s].y += xv.y * t; }
                else acc[s] += xv * t;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < FMAX; s++) {
        long long i = r * F + s;
        if (s < F && i < new_points) reinterpret_cast<X*>(y_)[i] = acc[s];
    }
}

// generic-factor variant of the same polyphase kernel for F > 8: one thread per output
template <typename T, bool CPLX>
__global__ void interp_poly_generic_kernel(const void* __restrict__ x_, void* __restrict__ y_, const T* __restrict__ tab,
                                           long long N, long long new_points, int F, int L, long long scalar_len) {
    typedef typename CpxOf<T>::type C;
    typedef typename std::conditional<CPLX, C, T>::type X;
    const int J = 2 * L + 3;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= new_points) return;
    long long r = i / F; int s = (int)(i - r * F);
    bool interior = (i >= scalar_len) && (i < new_points - scalar_len);
    const T* t = tab + (interior ? 0 : F * J) + s * J;
    long long g = (r - L - 1) % N; if (g < 0) g += N;
    X acc = X();
    for (int j = 0; j < J; j++) {
        X xv = reinterpret_cast<const X*>(x_)[g];
        if constexpr (CPLX) { acc.x += xv.x * t[j]; acc.y += xv.y * t[j]; } else acc += xv * t[j];
        g++; if (g >= N) g -= N;
    }
    reinterpret_cast<X*>(y_)[i] = acc;
}

template <typename T>
int interp_poly(const void* x, void* y, const T* tab_dev, size_t N, size_t new_points, int F, int L, int is_complex,
                cudaStream_t st) {
    typedef typename CpxOf<T>::type C;
    const long long scalar_len = (long long)(2 * L + 1) * F;
    const int J = 2 * L + 3;
    if (F <= 8) {
        const size_t xs = is_complex ? sizeof(C) : sizeof(T);
        const size_t tab_pad = ((size_t)2 * F * J + 3) & ~(size_t)3;
        const size_t smem = tab_pad * sizeof(T) + (IP_THREADS + J) * xs;
        const long long rows = ((long long)new_points + F - 1) / F;
        const long long grid = (rows + IP_THREADS - 1) / IP_THREADS;
#define BDSP_IP(CP, FM)                                                                                              \
    do {                                                                                                             \
        if (smem > 48 * 1024) BDSP_CUDA_OK(cudaFuncSetAttribute(interp_poly_kernel<T, CP, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        interp_poly_kernel<T, CP, FM><<<(unsigned)grid, IP_THREADS, smem, st>>>(x, y, tab_dev, (long long)N, (long long)new_points, F, L, scalar_len); \
    } while (0)
        if (is_complex) { if (F <= 4) BDSP_IP(true, 4); else BDSP_IP(true, 8); }
        else { if (F <= 4) BDSP_IP(false, 4); else BDSP_IP(false, 8); }
#undef BDSP_IP
    } else {
        const long long grid = ((long long)new_points + 255) / 256;
        if (is_complex) interp_poly_generic_kernel<T, true><<<(unsigned)grid, 256, 0, st>>>(x, y, tab_dev, (long long)N, (long long)new_points, F, L, scalar_len);
        else interp_poly_generic_kernel<T, false><<<(unsigned)grid, 256, 0, st>>>(x, y, tab_dev, (long long)N, (long long)new_points, F, L, scalar_len);
    }
    BDSP_LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Non-integer factor: interpolate_priv_scalar (interpolation.rs:92-131).  Taps are evaluated per
// output on the device in precision T for the two built-in impulse responses
// (conv_types.rs:406-423 RaisedCosine, :479-487 Sinc).
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T pi_t() { return (T)3.14159265358979323846; }

template <typename T> __device__ __forceinline__ T sinc_calc(T x) {
    if (x == (T)0) return (T)1;
    T pi_x = pi_t<T>() * x;
    return sin(pi_x) / pi_x;
}

template <typename T> __device__ __forceinline__ T rc_calc(T x, T rolloff) {
    if (x == (T)0) return (T)1;
    const T one = (T)1, two = (T)2, pi = pi_t<T>();
    const T four = two * two;
    if (fabs(x) == one / (two * rolloff)) {
        T arg = pi / two / rolloff;
        return sin(arg) / arg * pi / four;
    }
    T pi_x = pi * x;
    T arg = two * rolloff * x;
    return sin(pi_x) * cos(pi_x * rolloff) / pi_x / (one - (arg * arg));
}

template <typename T, bool CPLX>
__global__ void interp_frac_kernel(const void* __restrict__ x_, void* __restrict__ y_, long long N, long long new_points,
                                   T factor, T delay, int L, int kind, T rolloff) {
    typedef typename CpxOf<T>::type C;
    typedef typename std::conditional<CPLX, C, T>::type X;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= new_points) return;
    const T center = (T)i / factor;
    const T rounded = floor(center);
    long long r = (long long)rounded;
    T j = -(T)L - (center - rounded) + delay;
    long long g = (r - L) % N; if (g < 0) g += N;
    X acc = X();
    for (int k = 0; k < 2 * L + 1; k++) {
        const T t = kind == 0 ? sinc_calc<T>(j) : rc_calc<T>(j, rolloff);
        X xv = reinterpret_cast<const X*>(x_)[g];
        if constexpr (CPLX) { acc.x += xv.x * t; acc.y += xv.y * t; } else acc += xv * t;
        j = j + (T)1;
        g++; if (g >= N) g -= N;
    }
    reinterpret_cast<X*>(y_)[i] = acc;
}

template <typename T>
int interp_frac(const void* x, void* y, size_t N, size_t new_points, double factor, double delay, int L, int kind,
                double rolloff, int is_complex, cudaStream_t st) {
    const long long grid = ((long long)new_points + 255) / 256;
    if (is_complex) interp_frac_kernel<T, true><<<(unsigned)grid, 256, 0, st>>>(x, y, (long long)N, (long long)new_points, (T)factor, (T)delay, L, kind, (T)rolloff);
    else interp_frac_kernel<T, false><<<(unsigned)grid, 256, 0, st>>>(x, y, (long long)N, (long long)new_points, (T)factor, (T)delay, L, kind, (T)rolloff);
    BDSP_LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------
// interpolate_lin (real_interpolation.rs:33-71): every operation individually rounded in T.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float lin_eval(float i, float F, float d, const float* __restrict__ x, long long n) {
    float p = __fadd_rn(__fdiv_rn(i, F), d);
    float bf = floorf(p);
    long long b = (long long)bf;
    if (b < 0) b = 0;                // the reference panics on an out-of-range index; clamp instead
    if (b > n - 2) b = n - 2;
    float y0 = x[b], y1 = x[b + 1];
    return __fadd_rn(y0, __fmul_rn(__fsub_rn(y1, y0), __fsub_rn(p, bf)));
}
__device__ __forceinline__ double lin_eval(double i, double F, double d, const double* __restrict__ x, long long n) {
    double p = __dadd_rn(__ddiv_rn(i, F), d);
    double bf = floor(p);
    long long b = (long long)bf;
    if (b < 0) b = 0;
    if (b > n - 2) b = n - 2;
    double y0 = x[b], y1 = x[b + 1];
    return __dadd_rn(y0, __dmul_rn(__dsub_rn(y1, y0), __dsub_rn(p, bf)));
}

// The reference counts the output index in precision T by repeated `+ 1.0` (real_interpolation.rs:53,65);
// in f32 that counter stops advancing at 2^24 (Q7).  The same saturation is applied here so that the
// result is bit-identical to the reference for every length (and the read index stays in bounds).
__device__ __forceinline__ float counter_value(long long i, float) { return i >= 16777216ll ? 16777216.0f : (float)i; }
__device__ __forceinline__ double counter_value(long long i, double) { return i >= 9007199254740992ll ? 9007199254740992.0 : (double)i; }

template <typename T>
__global__ void interp_lin_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, long long dest_len, T F, T d) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < dest_len; i += stride) {
        if (i == dest_len - 1) y[i] = x[n - 1];
        else y[i] = lin_eval(counter_value(i, (T)0), F, d, x, n);
    }
}

template <typename T>
int interp_lin(const void* x, void* y, size_t n, size_t dest_len, double factor, double delay, cudaStream_t st) {
    long long grid = ((long long)dest_len + 255) / 256;
    const long long cap = (long long)sm_count() * 32;
    if (grid > cap) grid = cap;
    interp_lin_kernel<T><<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), (long long)n,
                                                       (long long)dest_len, (T)factor, (T)delay);
    BDSP_LAUNCHED();
    return 0;
}

#define BDSP_INST(T)                                                                                                  \
    template int interp_poly<T>(const void*, void*, const T*, size_t, size_t, int, int, int, cudaStream_t);             \
    template int interp_frac<T>(const void*, void*, size_t, size_t, double, double, int, int, double, int, cudaStream_t); \
    template int interp_lin<T>(const void*, void*, size_t, size_t, double, double, cudaStream_t);
BDSP_INST(float)
BDSP_INST(double)
#undef BDSP_INST

}  // namespace bdsp
