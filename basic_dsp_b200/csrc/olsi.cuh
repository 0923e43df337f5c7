// Pieces shared by the fused overlap-save kernels on interleaved-complex FP32x2 arithmetic (ols4096i.cu, ols8192i.cu).
#pragma once
#include "cxmath.cuh"

namespace bdsp {
using namespace cx;

// block inputs are used once: OI_X_LOAD = 1 (L1::no_allocate) / 2 (ld.global.cg, L2 only) keep them from displacing the
// spectrum and the twiddles in L1
#ifndef OI_X_LOAD
#define OI_X_LOAD 0
#endif
__device__ __forceinline__ float4 ld_x4(const float2* p) {
#if OI_X_LOAD == 1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#elif OI_X_LOAD == 2
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return __ldg(reinterpret_cast<const float4*>(p));
#endif
}
// stride-16 stage twiddles W256^{k c}, k = 1..15: 0 = fifteen 128-bit table loads per stage (4-5 L1 wavefronts each:
// 16 % of the kernel's L1 data-pipe traffic, which is its saturated unit), 1 = six loads (k = 1, 2, 3, 4, 8, 12) +
// nine products, 2 = one load + the power scheme of the stride-256 stages
// measurement only (results become wrong): bit 0 = no F3|H|I3 arithmetic, 1 = no stride-16 stages' arithmetic, 2 = no
// stride-256 stages' arithmetic, 3 = no H fetch, 4 = no shared-memory traffic in the middle section
#ifndef OI_ABLATE
#define OI_ABLATE 0
#endif
#ifndef OI_TW2_MODE
#define OI_TW2_MODE 1
#endif
// v[s] *= W^{k(s)} (CONJ: conjugated), k(s) = SLOTMAP ? r16_k(s) : s, for the two columns held in v / u
// tw2[STRIDE * k]: float4 = W^{k c}, W^{k (c + 1)} of the thread's two columns
template <bool SLOTMAP, bool CONJ, int STRIDE>
__device__ __forceinline__ void oi_tw2(c2 (&v)[16], c2 (&u)[16], const float4* __restrict__ tw2) {
#if OI_TW2_MODE == 0
#pragma unroll
    for (int s = 1; s < 16; s++) {
        const float4 f = __ldg(tw2 + STRIDE * (SLOTMAP ? r16_k(s) : s));
        v[s] = CONJ ? mul_conj(v[s], make_float2(f.x, f.y)) : mul(v[s], make_float2(f.x, f.y));
        u[s] = CONJ ? mul_conj(u[s], make_float2(f.z, f.w)) : mul(u[s], make_float2(f.z, f.w));
    }
#elif OI_TW2_MODE == 1
    c2 Av[4], Bv[4], Au[4], Bu[4];
#pragma unroll
    for (int i = 1; i < 4; i++) {
        const float4 fa = __ldg(tw2 + STRIDE * i), fb = __ldg(tw2 + STRIDE * 4 * i);
        Av[i] = make_float2(fa.x, CONJ ? -fa.y : fa.y); Au[i] = make_float2(fa.z, CONJ ? -fa.w : fa.w);
        Bv[i] = make_float2(fb.x, CONJ ? -fb.y : fb.y); Bu[i] = make_float2(fb.z, CONJ ? -fb.w : fb.w);
    }
#pragma unroll
    for (int s = 1; s < 16; s++) {
        const int k = SLOTMAP ? r16_k(s) : s;
        const int a = k & 3, b = k >> 2;
        const c2 wv = b == 0 ? Av[a] : a == 0 ? Bv[b] : mul(Av[a], Bv[b]);
        const c2 wu = b == 0 ? Au[a] : a == 0 ? Bu[b] : mul(Au[a], Bu[b]);
        v[s] = mul(v[s], wv);
        u[s] = mul(u[s], wu);
    }
#else
    const float4 f = __ldg(tw2 + STRIDE);
    apply_twiddles<SLOTMAP>(v, make_float2(f.x, CONJ ? -f.y : f.y));
    apply_twiddles<SLOTMAP>(u, make_float2(f.z, CONJ ? -f.w : f.w));
#endif
}

// plan geometry shared by prepare and convolve (block length M): the delay d in {0, 1} samples that the plan folds
// into H makes the distance `shift` between a block position's input index and its output index even, so that with even
// block offsets every thread moves its two adjacent points with one 128-bit access on both sides.
// m_first: first block position whose circular-convolution value is valid; step: outputs per block.
static inline void olsi_geometry(size_t M, size_t L, int* d, int* shift, int* m_first, int* step) {
    const int cl = (int)(L - L / 2);
    *d = (cl - 1) & 1;
    *shift = cl - 1 + *d;
    int mf = (int)L - 1 + *d;
    if (mf & 1) mf++;
    *m_first = mf;
    *step = ((int)M - mf) & ~1;
}

}  // namespace bdsp
