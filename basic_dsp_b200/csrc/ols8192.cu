// Fused 8192-point overlap-save block: forward FFT, spectrum multiply and inverse FFT of one block in ONE kernel
// pass with the block resident in shared memory (the reference's overlap_discard, convolution.rs:304-461, whose
// blocks are 2 rustfft calls + a scalar multiply loop with the spectrum going through memory).
//
// Why a second block length next to ols4096: the reference derives fft_len from the tap count (convolution.rs:323-331);
// for the 1023-tap config a 4096-point block keeps 3072/4096 = 75 % of its points, an 8192-point block 7168/8192 =
// 87.5 %.  Per OUTPUT sample this kernel needs 14 % fewer shared-memory wavefronts and ~4 % fewer FP32x2 instructions
// than ols4096 - the two saturated units of that kernel (profiles/r1_ols4096_ncu.txt: LSU 78 %, FMA 66 %) - and it
// serves responses of up to 4094 taps.
//
// Structure (M = 8192 = 16 * 16 * 32, 256 threads, 32 points per thread and stage, planar shared memory, FP32x2 math):
//   F1  radix-16 DIF over stride 512, inputs straight from global memory            -> smem     | CTA barrier
//   F2  radix-16 DIF over stride 32 inside every 512-point row                      -> smem     | warp barrier
//   F3  radix-32 DIF over 32 contiguous points | * H (position order) | radix-32 DIT (registers) -> smem  | warp barrier
//   I2  radix-16 DIT over stride 32                                                 -> smem     | CTA barrier
//   I1  radix-16 DIT over stride 512, valid outputs straight to global memory
// Row k0 (512 points) is owned by half-warp k0 in F2, F3 and I2, so three of the five exchanges only need
// __syncwarp(): two CTA-wide barriers per block instead of the four of ols4096.
// Layout: float index of block position p = 512*row + 32*g + n0 is 512*row + 32*g + (n0 ^ 4*(g & 7)): every
// 64-bit access of a half-warp covers one 128-byte line, and the eight 128-bit accesses of a quarter-warp in F3
// (consecutive g) fall into eight different 16-byte bank windows.  No padding: exactly 64 KB per CTA.
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "conv.cuh"
#include "ols4096.cuh"

namespace bdsp {

using namespace ols16;

#define O8_M 8192
#define O8_THREADS 256
#ifndef O8_MIN_CTAS
#define O8_MIN_CTAS 2
#endif
#ifndef O8_NT
#define O8_NT 256        // threads per CTA (256: one virtual thread per thread, two CTAs per SM; 128: two, three CTAs per SM;
#endif                   // measured on B200, 64 x 2^20 points, 1023 taps: 0.360 ms vs 0.393 ms)
#if 0
#endif
#ifndef O8_TW2_TEX
#define O8_TW2_TEX 0     // stride-32 stage twiddles through the texture path instead of LDG
#endif
#ifndef O8_H_TEX
#define O8_H_TEX 1
#endif
#ifndef O8_TW2_SMEM
#define O8_TW2_SMEM 1    // stride-32 stage twiddles staged in shared memory (4 KB per CTA): loaded where they are used, so the
#endif                   // compiler cannot hoist them above the barriers (LDG version: 52 registers of prefetched twiddles, spills)
#ifndef O8_TW2_COMPUTE
#define O8_TW2_COMPUTE 0  // stride-32 stage twiddles as powers of one root per thread (no table loads, +112 FP32x2 instructions)
#endif
#ifndef O8_TW2_DEP
#define O8_TW2_DEP 0      // LDG twiddles whose address depends on a shared-memory word read after the barrier (no hoisting)
#endif
#define O8_SMEM_BYTES ((2 * O8_M + (O8_TW2_SMEM ? 1024 : O8_TW2_DEP ? 4 : 0)) * sizeof(float))

// twiddle table (floats):
//   [0,512) Re W8192^c, [512,1024) Im W8192^c                                   (base roots of the stride-512 stages)
//   float4 index 256 + 16*k + n/2 = {Re W512^{k n}, Re W512^{k (n+1)}, Im .., Im ..}, n even in [0,32), k in [0,16)
#define O8_TW_FLOATS (1024 + 1024)
#define O8_TW2_F4 256

__device__ __forceinline__ void sts64s(float* p, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y) : "memory");
}

// cos / sin (pi j / 16), j = 0..15: W32^j = C32[j] - i S32[j]
#define O8_C1 0.98078528040323044913f
#define O8_C2 0.92387953251128675613f
#define O8_C3 0.83146961230254523708f
#define O8_C4 0.70710678118654752440f
#define O8_C5 0.55557023301960222474f
#define O8_C6 0.38268343236508977173f
#define O8_C7 0.19509032201612826785f

// 32 CONTIGUOUS points, packed over adjacent points: P[m] = {a[2m], a[2m+1]}.  DIF, natural in -> bit-reversed out
// (slot j holds frequency bitrev5(j)): the span-16 level below, then fft16_dif on P[0..7] and P[8..15].
__device__ __forceinline__ void fft32_span16_fwd(cp (&P)[16]) {
    // twiddle W32^j, j = 2m, 2m+1
    const pk tr[8] = {make_float2(1.f, O8_C1), make_float2(O8_C2, O8_C3), make_float2(O8_C4, O8_C5), make_float2(O8_C6, O8_C7),
                      make_float2(0.f, -O8_C7), make_float2(-O8_C6, -O8_C5), make_float2(-O8_C4, -O8_C3), make_float2(-O8_C2, -O8_C1)};
    const pk ti[8] = {make_float2(0.f, -O8_C7), make_float2(-O8_C6, -O8_C5), make_float2(-O8_C4, -O8_C3), make_float2(-O8_C2, -O8_C1),
                      make_float2(-1.f, -O8_C1), make_float2(-O8_C2, -O8_C3), make_float2(-O8_C4, -O8_C5), make_float2(-O8_C6, -O8_C7)};
#pragma unroll
    for (int m = 0; m < 8; m++) {
        cp u = cadd(P[m], P[m + 8]);
        cp d = csub(P[m], P[m + 8]);
        P[m] = u;
        P[m + 8] = cmulc(d, tr[m], ti[m]);
    }
}

// inverse (DIT, bit-reversed in -> natural out): fft16_dit_inv on both halves, then this level
__device__ __forceinline__ void fft32_span16_inv(cp (&P)[16]) {
    const pk tr[8] = {make_float2(1.f, O8_C1), make_float2(O8_C2, O8_C3), make_float2(O8_C4, O8_C5), make_float2(O8_C6, O8_C7),
                      make_float2(0.f, -O8_C7), make_float2(-O8_C6, -O8_C5), make_float2(-O8_C4, -O8_C3), make_float2(-O8_C2, -O8_C1)};
    const pk ti[8] = {make_float2(0.f, O8_C7), make_float2(O8_C6, O8_C5), make_float2(O8_C4, O8_C3), make_float2(O8_C2, O8_C1),
                      make_float2(1.f, O8_C1), make_float2(O8_C2, O8_C3), make_float2(O8_C4, O8_C5), make_float2(O8_C6, O8_C7)};
#pragma unroll
    for (int m = 0; m < 8; m++) {
        cp t = cmulc(P[m + 8], tr[m], ti[m]);
        cp u = P[m];
        P[m] = cadd(u, t);
        P[m + 8] = csub(u, t);
    }
}

// special-register reads the compiler cannot merge with earlier ones (see I1)
__device__ __forceinline__ int fresh_tid() { int v; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(v)); return v; }
__device__ __forceinline__ int fresh_ctaid() { int v; asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(v)); return v; }

__host__ __device__ __forceinline__ int bitrev5(int j) {
    return ((j & 1) << 4) | ((j & 2) << 2) | (j & 4) | ((j & 8) >> 2) | ((j & 16) >> 4);
}
// frequency index held at block position p after F1, F2, F3
__host__ __device__ __forceinline__ int o8_freq_of_pos(int p) {
    const int k0 = p >> 9, k1 = (p >> 5) & 15, jj = p & 31;
    return k0 + 16 * k1 + 256 * bitrev5(jj);
}

// NT threads per CTA (256 or 128); every thread works on IT = 256/NT "virtual threads" tv = t + NT*it per stage.
// NT = 128: three CTAs per SM (three independently phased CTAs per scheduler instead of two) and 168 registers per
// thread, which lets the compiler issue the loads of the second virtual thread under the arithmetic of the first.
template <bool ALIGNED, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? O8_MIN_CTAS : 3)
ols8192_kernel(const float2* __restrict__ x, float2* __restrict__ y, int N, int m_first, int step, int shift,
               int blocks_per_vec, const float* __restrict__ Hre, const float* __restrict__ Him,
               const float* __restrict__ tw, cudaTextureObject_t htex, cudaTextureObject_t twtex) {
    constexpr int IT = O8_THREADS / NT;
    extern __shared__ __align__(16) float o8_smem[];
    float* sre = o8_smem;
    float* sim = o8_smem + O8_M;
    const int t = threadIdx.x;
    const int vec = blockIdx.x / blocks_per_vec;
    const int blk = blockIdx.x - vec * blocks_per_vec;
    const int i0 = blk * step;
    const float2* xr = x + (size_t)vec * (size_t)N;
    const int l16 = t & 15;                   // lane inside the half-warp; virtual thread tv: half-warp (= row in F2/F3/I2) tv >> 4
#if O8_TW2_DEP && !O8_TW2_SMEM
    if (t == 0) reinterpret_cast<volatile int*>(o8_smem + 2 * O8_M)[0] = 0;
#endif
#if O8_TW2_SMEM
    float4* s_tw2 = reinterpret_cast<float4*>(o8_smem + 2 * O8_M);
#pragma unroll
    for (int it = 0; it < IT; it++)           // visible after the first CTA barrier
        s_tw2[t + NT * it] = __ldg(reinterpret_cast<const float4*>(tw) + O8_TW2_F4 + t + NT * it);
#endif

    // ------------------------------------------------------------------ F1: stride 512, from global
    {
        // block position p holds x[(p0 + p) mod N], p0 = i0 + shift - m_first  (p0 > -8192, even)
        const int p0 = i0 + shift - m_first;
        const bool inside = p0 >= 0 && p0 + O8_M <= N;   // block-uniform: no wrap-around inside this block
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int tv = t + NT * it, hw = tv >> 4;
            const int c = 2 * tv;
            cp v[16];
            if (inside) {
                const float2* px = xr + p0 + c;
#pragma unroll
                for (int a = 0; a < 16; a++) {
                    if (ALIGNED) {
                        const float4 ab = ldg_x4(reinterpret_cast<const float4*>(px + 512 * a));
                        v[a].re = make_float2(ab.x, ab.z);
                        v[a].im = make_float2(ab.y, ab.w);
                    } else {
                        const float2 p = __ldg(px + 512 * a);
                        const float2 q = __ldg(px + 512 * a + 1);
                        v[a].re = make_float2(p.x, q.x);
                        v[a].im = make_float2(p.y, q.y);
                    }
                }
            } else {                             // first / last block of a vector: circular indexing
                int idx = (p0 + c) % N;
                if (idx < 0) idx += N;
                const int adv = 512 % N;
#pragma unroll
                for (int a = 0; a < 16; a++) {
                    int i1 = idx + 1; if (i1 >= N) i1 -= N;
                    const float2 p = __ldg(&xr[idx]);
                    const float2 q = __ldg(&xr[i1]);
                    v[a].re = make_float2(p.x, q.x);
                    v[a].im = make_float2(p.y, q.y);
                    idx += adv; if (idx >= N) idx -= N;
                }
            }
            r16<false>(v);
            cp w1;
            w1.re = __ldg(reinterpret_cast<const float2*>(tw + c));
            w1.im = __ldg(reinterpret_cast<const float2*>(tw + 512 + c));
            apply_twiddles<true>(v, w1);
            // column pair c = 32*g + n0: g = hw, n0 = 2*l16
            const int off = 32 * hw + ((2 * l16) ^ (4 * (hw & 7)));
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const int k0 = r16_k(s);
                *reinterpret_cast<float2*>(&sre[512 * k0 + off]) = v[s].re;
                *reinterpret_cast<float2*>(&sim[512 * k0 + off]) = v[s].im;
            }
        }
    }
    __syncthreads();
    const int n0 = 2 * l16;
#if O8_TW2_SMEM
#define O8_TW2(k1) s_tw2[16 * (k1) + l16]
#elif O8_TW2_TEX
#define O8_TW2(k1) tex1Dfetch<float4>(twtex, O8_TW2_F4 + 16 * (k1) + l16)
#else
#if O8_TW2_DEP
    const float4* tw2 = reinterpret_cast<const float4*>(tw) + O8_TW2_F4 + l16 + reinterpret_cast<volatile int*>(o8_smem + 2 * O8_M)[0];
#else
    const float4* tw2 = reinterpret_cast<const float4*>(tw) + O8_TW2_F4 + l16;
#endif
#define O8_TW2(k1) __ldg(tw2 + 16 * (k1))
#endif
    // ------------------------------------------------------------------ F2: stride 32 inside row hw
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int hw = (t + NT * it) >> 4;
        float* rre = sre + 512 * hw;
        float* rim = sim + 512 * hw;
        cp v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const int a = 32 * n1 + (n0 ^ (4 * (n1 & 7)));
            v[n1].re = *reinterpret_cast<const float2*>(&rre[a]);
            v[n1].im = *reinterpret_cast<const float2*>(&rim[a]);
        }
        r16<false>(v);
#if O8_TW2_COMPUTE
        {
            const float4 f = O8_TW2(1);
            cp w1;
            w1.re = make_float2(f.x, f.y);
            w1.im = make_float2(f.z, f.w);
            apply_twiddles<true>(v, w1);
        }
#else
#pragma unroll
        for (int s = 1; s < 16; s++) {
            const int k1 = r16_k(s);
            const float4 f = O8_TW2(k1);
            cp w;
            w.re = make_float2(f.x, f.y);
            w.im = make_float2(f.z, f.w);
            v[s] = cmul(v[s], w);
        }
#endif
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int k1 = r16_k(s);
            const int a = 32 * k1 + (n0 ^ (4 * (k1 & 7)));
            *reinterpret_cast<float2*>(&rre[a]) = v[s].re;
            *reinterpret_cast<float2*>(&rim[a]) = v[s].im;
        }
    }
    __syncwarp();
    // ------------------------------------------------------------------ F3 | *H | I3 on the 32 contiguous points (hw, k1 = l16)
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int tv = t + NT * it, hw = tv >> 4;
        float* rre = sre + 512 * hw;
        float* rim = sim + 512 * hw;
        const int base = 32 * l16;
        const int sw = 4 * (l16 & 7);
        cp P[16];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int a = base + ((4 * q) ^ sw);
            const float4 fr = *reinterpret_cast<const float4*>(&rre[a]);
            const float4 fi = *reinterpret_cast<const float4*>(&rim[a]);
            P[2 * q].re = make_float2(fr.x, fr.y); P[2 * q + 1].re = make_float2(fr.z, fr.w);
            P[2 * q].im = make_float2(fi.x, fi.y); P[2 * q + 1].im = make_float2(fi.z, fi.w);
        }
        // radix-32 = one radix-2 level over span 16 + two independent 16-point transforms: each half goes through
        // F3 | *H | I3 on its own
        fft32_span16_fwd(P);
#pragma unroll
        for (int hf = 0; hf < 2; hf++) {
            cp(&Q)[8] = *reinterpret_cast<cp(*)[8]>(&P[8 * hf]);
            fft16_dif<false>(Q);
#pragma unroll
            for (int qq = 0; qq < 4; qq++) {
                const int q = 4 * hf + qq;
                // plan layout: [q][virtual thread][4] -> a warp reads 512 contiguous bytes per request
#if O8_H_TEX
                const float4 hr = tex1Dfetch<float4>(htex, q * O8_THREADS + tv);
                const float4 hi = tex1Dfetch<float4>(htex, 2048 + q * O8_THREADS + tv);
#else
                const float4 hr = __ldg(reinterpret_cast<const float4*>(Hre) + q * O8_THREADS + tv);
                const float4 hi = __ldg(reinterpret_cast<const float4*>(Him) + q * O8_THREADS + tv);
#endif
                cp h0, h1;
                h0.re = make_float2(hr.x, hr.y); h0.im = make_float2(hi.x, hi.y);
                h1.re = make_float2(hr.z, hr.w); h1.im = make_float2(hi.z, hi.w);
                Q[2 * qq] = cmul(Q[2 * qq], h0);
                Q[2 * qq + 1] = cmul(Q[2 * qq + 1], h1);
            }
            fft16_dit_inv(Q);
        }
        fft32_span16_inv(P);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int a = base + ((4 * q) ^ sw);
            sts64s(&rre[a], P[2 * q].re);
            sts64s(&rre[a + 2], P[2 * q + 1].re);
            sts64s(&rim[a], P[2 * q].im);
            sts64s(&rim[a + 2], P[2 * q + 1].im);
        }
    }
    __syncwarp();
    // ------------------------------------------------------------------ I2: stride 32 (DIT: twiddle first)
#if !O8_TW2_TEX && !O8_TW2_SMEM
    // opaque copy of the table pointer: without it the compiler keeps the 15 twiddles F2 loaded (60 registers) alive
    // across F3 and spills them
#if O8_TW2_DEP
    const float4* tw2i = reinterpret_cast<const float4*>(tw) + O8_TW2_F4 + l16 + reinterpret_cast<volatile int*>(o8_smem + 2 * O8_M)[0];
#else
    const float4* tw2i = tw2;
    asm volatile("" : "+l"(tw2i));
#endif
#undef O8_TW2
#define O8_TW2(k1) __ldg(tw2i + 16 * (k1))
#endif
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int hw = (t + NT * it) >> 4;
        float* rre = sre + 512 * hw;
        float* rim = sim + 512 * hw;
        cp v[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            const int a = 32 * k1 + (n0 ^ (4 * (k1 & 7)));
            v[k1].re = *reinterpret_cast<const float2*>(&rre[a]);
            v[k1].im = *reinterpret_cast<const float2*>(&rim[a]);
        }
#if O8_TW2_COMPUTE
        {
            const float4 f = O8_TW2(1);
            cp w1;
            w1.re = make_float2(f.x, f.y);
            w1.im = pneg(make_float2(f.z, f.w));
            apply_twiddles<false>(v, w1);
        }
#else
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) {
            const float4 f = O8_TW2(k1);
            cp w;
            w.re = make_float2(f.x, f.y);
            w.im = make_float2(f.z, f.w);
            v[k1] = cmul_conj(v[k1], w);
        }
#endif
        r16<true>(v);
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int n1 = r16_k(s);
            const int a = 32 * n1 + (n0 ^ (4 * (n1 & 7)));
            *reinterpret_cast<float2*>(&rre[a]) = v[s].re;
            *reinterpret_cast<float2*>(&rim[a]) = v[s].im;
        }
    }
    __syncthreads();
    // ------------------------------------------------------------------ I1: stride 512, valid outputs to global
    {
        // thread / block geometry is re-derived from the special registers here: nothing of it stays live across the
        // register-hungry middle section (F2 .. I2)
        const int t = fresh_tid();
        const int bid = fresh_ctaid();
        const int vec = bid / blocks_per_vec;
        const int i0 = (bid - vec * blocks_per_vec) * step;
        float2* yr = y + (size_t)vec * (size_t)N;
        float* sre = o8_smem;
        float* sim = o8_smem + O8_M;
        int mhi = step;                                    // exclusive bound on m
        if (i0 + step > N) mhi = N - i0;                   // last block of the vector
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int tv = t + NT * it, hw = tv >> 4, l16 = tv & 15;
            const int c = 2 * tv;
            const int off = 32 * hw + ((2 * l16) ^ (4 * (hw & 7)));
            cp v[16];
#pragma unroll
            for (int k0 = 0; k0 < 16; k0++) {
                v[k0].re = *reinterpret_cast<const float2*>(&sre[512 * k0 + off]);
                v[k0].im = *reinterpret_cast<const float2*>(&sim[512 * k0 + off]);
            }
            cp w1;
            w1.re = __ldg(reinterpret_cast<const float2*>(tw + c));
            w1.im = pneg(__ldg(reinterpret_cast<const float2*>(tw + 512 + c)));
            apply_twiddles<false>(v, w1);
            r16<true>(v);
            // output i = i0 + m, m = c + 512*a - m_first in [0, step) and i < N
            const int mlo = c - m_first;                       // m for a = 0
            float2* py = yr + i0 + mlo;
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const int a = r16_k(s);
                const int m = mlo + 512 * a;
                if (ALIGNED) {
                    if (m >= 0 && m < mhi)
                        *reinterpret_cast<float4*>(py + 512 * a) = make_float4(v[s].re.x, v[s].im.x, v[s].re.y, v[s].im.y);
                } else {
                    if (m >= 0 && m < mhi) py[512 * a] = make_float2(v[s].re.x, v[s].im.x);
                    if (m + 1 >= 0 && m + 1 < mhi) py[512 * a + 1] = make_float2(v[s].re.y, v[s].im.y);
                }
            }
        }
    }
#undef O8_TW2
}

// Hpos (planar, kernel layout) <- Hs (interleaved, natural order, already scaled by 1/M), delayed by d samples:
// H_d[k] = H[k] * exp(-2 pi i k d / M).  Kernel layout: the value of block position p = 32*t + 4*q + e
// (t = 16*row + g) is stored at (q*256 + t)*4 + e.
__global__ void ols8192_permute_h_kernel(const float2* __restrict__ Hs, float* __restrict__ Hre, float* __restrict__ Him, int d) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= O8_M) return;
    const int k = o8_freq_of_pos(p);
    float2 h = Hs[k];
    if (d) h = cmul(h, unit_root<float>((unsigned long long)k * (unsigned long long)d, O8_M, -1));
    const int t = p >> 5, q = (p >> 2) & 7, e = p & 3;
    const int dst = (q * O8_THREADS + t) * 4 + e;
    Hre[dst] = h.x;
    Him[dst] = h.y;
}

namespace {
std::mutex g_o8_mu;
struct O8Dev { float* tw = nullptr; cudaTextureObject_t twtex = 0; bool attr = false; };
std::map<int, O8Dev> g_o8;

int o8_device(O8Dev** out) {
    int d = 0;
    BDSP_CUDA_OK(cudaGetDevice(&d));
    std::lock_guard<std::mutex> lk(g_o8_mu);
    O8Dev& s = g_o8[d];
    if (!s.tw) {
        std::vector<float> h(O8_TW_FLOATS);
        const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
        for (int c = 0; c < 512; c++) {
            h[c] = (float)cosl(-tau * (long double)c / 8192.0L);
            h[512 + c] = (float)sinl(-tau * (long double)c / 8192.0L);
        }
        for (int k = 0; k < 16; k++)
            for (int n = 0; n < 32; n++) {
                const long double a = -tau * (long double)((k * n) % 512) / 512.0L;
                float* e = &h[4 * (O8_TW2_F4 + 16 * k + n / 2)];
                e[n & 1] = (float)cosl(a);
                e[2 + (n & 1)] = (float)sinl(a);
            }
        float* dev = nullptr;
        BDSP_CUDA_OK(cudaMalloc(&dev, O8_TW_FLOATS * sizeof(float)));
        BDSP_CUDA_OK(cudaMemcpy(dev, h.data(), O8_TW_FLOATS * sizeof(float), cudaMemcpyHostToDevice));
        s.tw = dev;
#if O8_TW2_TEX
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = dev;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = O8_TW_FLOATS * sizeof(float);
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        BDSP_CUDA_OK(cudaCreateTextureObject(&s.twtex, &rd, &td, nullptr));
#endif
    }
    if (!s.attr) {
        BDSP_CUDA_OK(cudaFuncSetAttribute(ols8192_kernel<true, O8_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)O8_SMEM_BYTES));
        BDSP_CUDA_OK(cudaFuncSetAttribute(ols8192_kernel<false, O8_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)O8_SMEM_BYTES));
        s.attr = true;
    }
    *out = &s;
    return 0;
}
}  // namespace

bool ols8192_applicable(size_t N, size_t L) {
    // one wrap at most per strided load needs N >= 512; 32-bit row indices
    return L >= 2 && L <= O8_M / 2 - 2 && N >= O8_M && N < (1ull << 30);
}

// plan geometry shared by prepare and convolve: delay d makes the input->output index distance even
static inline void ols8192_geometry(size_t L, int* d, int* shift, int* m_first, int* step) {
    const int cl = (int)(L - L / 2);
    *d = (cl - 1) & 1;
    *shift = cl - 1 + *d;
    int mf = (int)L - 1 + *d;      // first block position whose circular convolution value is valid
    if (mf & 1) mf++;
    *m_first = mf;
    *step = (O8_M - mf) & ~1;
}

// Hpos: 2*8192 floats (re plane, im plane) <- Hs = FFT_8192(pad(h)) / 8192 (natural order)
int ols8192_prepare(const void* Hs, void* Hpos, size_t L, cudaStream_t st) {
    int d, shift, m_first, step;
    ols8192_geometry(L, &d, &shift, &m_first, &step);
    float* hp = reinterpret_cast<float*>(Hpos);
    ols8192_permute_h_kernel<<<O8_M / 256, 256, 0, st>>>(reinterpret_cast<const float2*>(Hs), hp, hp + O8_M, d);
    BDSP_LAUNCHED();
    return 0;
}

long long ols8192_blocks(size_t N, size_t L) {
    int d, shift, m_first, step;
    ols8192_geometry(L, &d, &shift, &m_first, &step);
    return ((long long)N + step - 1) / step;
}

int ols8192_convolve(const void* x, void* y, size_t N, size_t batch, size_t L, const void* Hpos, cudaTextureObject_t htex,
                     cudaStream_t st) {
    if (x == y) { set_last_error("ols8192_convolve: in-place operation is not supported"); return -3; }
    int d, shift, m_first, step;
    ols8192_geometry(L, &d, &shift, &m_first, &step);
    const long long bpv = ((long long)N + step - 1) / step;
    const long long grid = bpv * (long long)batch;
    if (grid > 0x7fffffffll) { set_last_error("ols8192_convolve: grid too large"); return -2; }
    O8Dev* dev = nullptr;
    int rc = o8_device(&dev);
    if (rc) return rc;
    const float* hp = reinterpret_cast<const float*>(Hpos);
    const bool aligned = (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const size_t smem = O8_SMEM_BYTES;
    if (aligned)
        ols8192_kernel<true, O8_NT><<<(unsigned)grid, O8_NT, smem, st>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(y), (int)N,
                                                                       m_first, step, shift, (int)bpv, hp, hp + O8_M, dev->tw, htex, dev->twtex);
    else
        ols8192_kernel<false, O8_NT><<<(unsigned)grid, O8_NT, smem, st>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(y), (int)N,
                                                                        m_first, step, shift, (int)bpv, hp, hp + O8_M, dev->tw, htex, dev->twtex);
    BDSP_LAUNCHED();
    return 0;
}

}  // namespace bdsp
