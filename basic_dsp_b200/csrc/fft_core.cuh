// Shared-memory Stockham FFT building blocks (power-of-two lengths), sm_100a.
//
// A CTA transforms `nfft` independent sequences of `n` complex points that live in shared memory
// (padded layout, see spad()).  Every stage reads R points per butterfly into registers, applies
// the stage twiddles, runs an in-register radix-R DIT transform and writes the R results back in
// Stockham (auto-sort) order, so the output is in natural order without a bit-reversal pass.
// Radices are chosen 16,16,...,{8,4,2}: the large radices come first, which keeps the strided
// Stockham writes of the early stages (stride Ns < 16) conflict free under the 1-in-16 padding.
#pragma once
#include "common.cuh"

namespace bdsp {

// log2 of the master twiddle table length: W[i] = exp(-2*pi*i*i/2^14), i in [0, 2^14)
#define BDSP_TW_LOG2 14
#define BDSP_TW_LEN (1 << BDSP_TW_LOG2)

// padded shared-memory index: one spare complex slot per 16 (8-byte banks: 16 lanes conflict free)
__device__ __forceinline__ int spad(int i) { return i + (i >> 4); }
static inline size_t spad_host(size_t n) { return n + (n >> 4) + 1; }

// ---- constants exp(-2 pi i k / R) for the in-register transforms --------------------------------
template <typename T, int R, int K> struct RootConst;  // only the non-trivial radix-16 roots are needed
#define BDSP_ROOT(R_, K_, RE_, IM_)                                         \
    template <typename T> struct RootConst<T, R_, K_> {                     \
        static __device__ __forceinline__ T re() { return (T)(RE_); }       \
        static __device__ __forceinline__ T im() { return (T)(IM_); }       \
    };
BDSP_ROOT(16, 1, 0.92387953251128675613, -0.38268343236508977173)
BDSP_ROOT(16, 3, 0.38268343236508977173, -0.92387953251128675613)
BDSP_ROOT(16, 5, -0.38268343236508977173, -0.92387953251128675613)
BDSP_ROOT(16, 7, -0.92387953251128675613, -0.38268343236508977173)
#undef BDSP_ROOT

// multiply by exp(-+2 pi i K/R) (INV flips the sign of the angle) with the trivial cases folded
template <typename T, int R, int K, bool INV, typename C> __device__ __forceinline__ C mul_root(C v) {
    if constexpr (K == 0) {
        return v;
    } else if constexpr (4 * K == R) {  // -i (forward) / +i (inverse)
        C r;
        if (!INV) { r.x = v.y; r.y = -v.x; } else { r.x = -v.y; r.y = v.x; }
        return r;
    } else if constexpr (8 * K == R) {  // (1 -+ i)/sqrt(2)
        const T h = (T)0.70710678118654752440;
        C r;
        if (!INV) { r.x = (v.x + v.y) * h; r.y = (v.y - v.x) * h; } else { r.x = (v.x - v.y) * h; r.y = (v.y + v.x) * h; }
        return r;
    } else if constexpr (8 * K == 3 * R) {  // (-1 -+ i)/sqrt(2)
        const T h = (T)0.70710678118654752440;
        C r;
        if (!INV) { r.x = (v.y - v.x) * h; r.y = -(v.x + v.y) * h; } else { r.x = -(v.x + v.y) * h; r.y = (v.x - v.y) * h; }
        return r;
    } else {
        const T wr = RootConst<T, R, K>::re();
        const T wi = INV ? -RootConst<T, R, K>::im() : RootConst<T, R, K>::im();
        C r;
        r.x = v.x * wr - v.y * wi;
        r.y = v.x * wi + v.y * wr;
        return r;
    }
}

// ---- in-register radix-R transform, natural order in, natural order out -------------------------
template <typename T, int R, bool INV> struct RegFFT;

template <typename T, bool INV> struct RegFFT<T, 1, INV> {
    typedef typename CpxOf<T>::type C;
    static __device__ __forceinline__ void run(C*) {}
};
template <typename T, bool INV> struct RegFFT<T, 2, INV> {
    typedef typename CpxOf<T>::type C;
    static __device__ __forceinline__ void run(C* v) {
        C a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <typename T, int R, int K, bool INV> struct CombineStep {
    typedef typename CpxOf<T>::type C;
    static __device__ __forceinline__ void run(C* v, const C* e, const C* o) {
        C t = mul_root<T, R, K, INV>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        CombineStep<T, R, K + 1, INV>::run(v, e, o);
    }
};
template <typename T, int R, bool INV> struct CombineStep<T, R, R / 2, INV> {
    typedef typename CpxOf<T>::type C;
    static __device__ __forceinline__ void run(C*, const C*, const C*) {}
};

template <typename T, int R, bool INV> struct RegFFT {
    typedef typename CpxOf<T>::type C;
    static __device__ __forceinline__ void run(C* v) {
        C e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        RegFFT<T, R / 2, INV>::run(e);
        RegFFT<T, R / 2, INV>::run(o);
        CombineStep<T, R, 0, INV>::run(v, e, o);
    }
};

// ---- one Stockham stage over data resident in (padded) shared memory ----------------------------
// s      : shared buffer holding nfft sequences, sequence f starts at element f*sstride (before padding)
// n      : points per sequence, Ns: product of the radices of the previous stages
// tw     : master twiddle table W_{2^14}^i (forward sign), in global memory
// Two __syncthreads() per stage (all reads, then all writes): the stage is in place.
template <typename T, int R, bool INV, int MAXB>
__device__ __forceinline__ void stockham_stage_strided(typename CpxOf<T>::type* s, int n, int nfft, int sstride, int Ns,
                                                       const typename CpxOf<T>::type* __restrict__ tw) {
    typedef typename CpxOf<T>::type C;
    const int bpf = n / R;            // butterflies per sequence (a power of two)
    const int lbpf = 31 - __clz(bpf);
    const int total = bpf * nfft;     // butterflies in this CTA
    // spad(a + b) = spad(a) + b + (b >> 4) when b is a multiple of 16: element strides that are multiples of 16 turn the
    // per-element padding into one constant stride
    const bool ld_fast = (bpf & 15) == 0;
    const int ld_stride = bpf + (bpf >> 4);
    const bool st_fast = (Ns & 15) == 0 || (Ns == 1 && (R == 8 || R == 16));
    const int st_stride = Ns == 1 ? 1 : Ns + (Ns >> 4);
    C v[MAXB][R];
    const int tw_scale = BDSP_TW_LEN / (Ns * R);
#pragma unroll
    for (int b = 0; b < MAXB; b++) {
        int w = threadIdx.x + b * blockDim.x;
        if (w < total) {
            int f = w >> lbpf, j = w & (bpf - 1);
            int k = j & (Ns - 1);
            const C* src = s;
            int base = f * sstride + j;
            if (ld_fast) {
                const C* p0 = src + spad(base);
#pragma unroll
                for (int r = 0; r < R; r++) v[b][r] = p0[r * ld_stride];
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) v[b][r] = src[spad(base + r * bpf)];
            }
            if (Ns > 1) {
                // one table load (W^k) per butterfly; the powers W^{k r} are products A_a * B_b, r = a + 4b
                // (depth <= 4).  R - 1 scattered table loads per butterfly made this stage L1-bound.
                C w1 = __ldg(&tw[k * tw_scale]);
                if (INV) w1.y = -w1.y;
                if (R <= 4) {
                    C w = w1;
#pragma unroll
                    for (int r = 1; r < R; r++) {
                        v[b][r] = cmul(v[b][r], w);
                        if (r + 1 < R) w = cmul(w, w1);
                    }
                } else {
                    C A[4], B[4];
                    A[1] = w1; A[2] = cmul(w1, w1); A[3] = cmul(A[2], w1);
                    B[1] = cmul(A[2], A[2]);
                    if (R > 8) { B[2] = cmul(B[1], B[1]); B[3] = cmul(B[2], B[1]); }
#pragma unroll
                    for (int r = 1; r < R; r++) {
                        const int a = r & 3, bb = r >> 2;
                        C w;
                        if (bb == 0) w = A[a];
                        else if (a == 0) w = B[bb];
                        else w = cmul(A[a], B[bb]);
                        v[b][r] = cmul(v[b][r], w);
                    }
                }
            }
            RegFFT<T, R, INV>::run(v[b]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < MAXB; b++) {
        int w = threadIdx.x + b * blockDim.x;
        if (w < total) {
            int f = w >> lbpf, j = w & (bpf - 1);
            int k = j & (Ns - 1);
            int obase = f * sstride + (j - k) * R + k;
            if (st_fast) {
                C* p0 = s + spad(obase);
#pragma unroll
                for (int r = 0; r < R; r++) p0[r * st_stride] = v[b][r];
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) s[spad(obase + r * Ns)] = v[b][r];
            }
        }
    }
    __syncthreads();
}

// Transform nfft sequences of n = 2^log2n points in shared memory.  Requires
// blockDim.x * MAXB16 >= nfft*n/16 (and the analogous bound for the remainder stage, see
// block_fft_threads()).  Caller must __syncthreads() after filling `s`.
template <typename T, bool INV>
__device__ __forceinline__ void block_fft(typename CpxOf<T>::type* s, int log2n, int nfft,
                                          const typename CpxOf<T>::type* __restrict__ tw) {
    const int n = 1 << log2n;
    int Ns = 1;
    int rem = log2n;
    while (rem >= 4) {
        stockham_stage_strided<T, 16, INV, 1>(s, n, nfft, n, Ns, tw);
        Ns <<= 4;
        rem -= 4;
    }
    if (rem == 3) stockham_stage_strided<T, 8, INV, 2>(s, n, nfft, n, Ns, tw);
    else if (rem == 2) stockham_stage_strided<T, 4, INV, 4>(s, n, nfft, n, Ns, tw);
    else if (rem == 1) stockham_stage_strided<T, 2, INV, 8>(s, n, nfft, n, Ns, tw);
}

// threads needed by block_fft for nfft sequences of n points (every stage fits its MAXB budget)
static inline int block_fft_threads(int n, int nfft) {
    int t = (n * nfft + 15) / 16;
    if (t < 32) t = 32;
    t = (t + 31) / 32 * 32;
    return t;
}

}  // namespace bdsp
