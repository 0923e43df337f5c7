// Fused 4096-point overlap-save kernel (see ols4096.cuh) + its plan helpers.
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "conv.cuh"
#include "ols4096.cuh"

namespace bdsp {

using namespace ols16;

// 64-bit shared store that the compiler cannot fuse with its neighbour into a 128-bit store
__device__ __forceinline__ void sts64(float* p, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y) : "memory");
}

// twiddle table layout (floats):
//   [0,256) Re W4096^c, [256,512) Im W4096^c                       (base roots of the stride-256 stages)
//   [512 + (8*k + n/2)*4 + {0,1,2,3}] = {Re W256^{k*n}, Re W256^{k*(n+1)}, Im W256^{k*n}, Im W256^{k*(n+1)}}, n even
//                                                                k, n in [0,16)  (stride-16 stages, one 128-bit load each)
//   [1024 + 256*k + c] Re W4096^{k*c}, [1024 + 4096 + 256*k + c] Im W4096^{k*c}  k in [0,16), c in [0,256)
#define OLS_TW_FLOATS (1024 + 8192)
#define OLS_TW2_RE 512
#define OLS_TW2_IM 768
#define OLS_TW1_RE 1024
#define OLS_TW1_IM (1024 + 4096)
#ifndef OLS_H_TEX
#define OLS_H_TEX 1   // measured: 0.3392 -> 0.3326 ms/step (LSU data pipe is the busiest unit, the texture pipe is idle)
#endif
#ifndef OLS_UNROLL_F3
#define OLS_UNROLL_F3 0
#endif
#ifndef OLS_MIN_CTAS
#define OLS_MIN_CTAS 5     // 96 registers, no spills.  With the warp-local middle section: 4 CTAs 0.3315 ms, 5 CTAs 0.3225 ms
#endif                     // per step of the bench (64 x 2^20 points, 1023 taps); 6 CTAs (80 registers) spill
#ifndef OLS_WARP_LOCAL
#define OLS_WARP_LOCAL 1   // F3 takes the two halves of the 256-point row its quarter-warp owns in F2 / I2: the F2 -> F3 and
#endif                     // I3 -> I2 exchanges then only need __syncwarp() (two CTA-wide barriers per block instead of four)
#ifndef OLS_COMPUTE_TW2
#define OLS_COMPUTE_TW2 0  // stride-16 stage twiddles as powers of one loaded root instead of 15 table loads
#endif
#ifndef OLS_TABLE_TW1
#define OLS_TABLE_TW1 0  // measured on B200: the table variant (L1-bound) is 10% slower than computing the powers
#endif

// ALIGNED: rows start on 16-byte boundaries (N even, 16-byte aligned base pointers); together with the
// even block offsets chosen by the plan every thread then moves its two adjacent points with one
// 128-bit access.  `shift` = cl - 1 + d is the (even) distance between a block position's input index
// and its output index.
template <bool ALIGNED>
__global__ void __launch_bounds__(OLS_THREADS, OLS_MIN_CTAS)
ols4096_kernel(const float2* __restrict__ x, float2* __restrict__ y, int N, int m_first, int step, int shift,
               int blocks_per_vec, const float* __restrict__ Hre, const float* __restrict__ Him,
               const float* __restrict__ tw, cudaTextureObject_t htex, cudaTextureObject_t xtex) {
    __shared__ __align__(16) float sre[OLS_PLANE];
    __shared__ __align__(16) float sim[OLS_PLANE];
    const int t = threadIdx.x;
    const int vec = blockIdx.x / blocks_per_vec;
    const int blk = blockIdx.x - vec * blocks_per_vec;
    const int i0 = blk * step;
    const float2* xr = x + (size_t)vec * (size_t)N;
    float2* yr = y + (size_t)vec * (size_t)N;

    cp v[16];
    // ------------------------------------------------------------------ F1: stride 256, from global
    {
        const int c = 2 * t;
        // block position m holds x[(p0 + m) mod N], p0 = i0 + shift - m_first  (p0 > -4096, even)
        const int p0 = i0 + shift - m_first;
        if (p0 >= 0 && p0 + OLS_M <= N) {   // block-uniform: no wrap-around inside this block
            const float2* px = xr + p0 + c;
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) {
                if (ALIGNED) {
                    const float4 ab = ldg_x4(reinterpret_cast<const float4*>(px + 256 * n2));
                    v[n2].re = make_float2(ab.x, ab.z);
                    v[n2].im = make_float2(ab.y, ab.w);
                } else {
                    const float2 a = __ldg(px + 256 * n2);
                    const float2 b = __ldg(px + 256 * n2 + 1);
                    v[n2].re = make_float2(a.x, b.x);
                    v[n2].im = make_float2(a.y, b.y);
                }
            }
        } else {                             // first / last block of a vector: circular indexing
            int idx = p0 + c;
            if (idx < 0) idx += N;
            if (idx >= N) idx -= N;
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) {
                int i1 = idx + 1; if (i1 >= N) i1 -= N;
                const float2 a = __ldg(&xr[idx]);
                const float2 b = __ldg(&xr[i1]);
                v[n2].re = make_float2(a.x, b.x);
                v[n2].im = make_float2(a.y, b.y);
                idx += 256; if (idx >= N) idx -= N;
            }
        }
        r16<false>(v);
#if OLS_TABLE_TW1
#pragma unroll
        for (int s = 1; s < 16; s++) {
            const int k0 = r16_k(s);
            cp w;
            w.re = __ldg(reinterpret_cast<const float2*>(tw + OLS_TW1_RE + 256 * k0 + c));
            w.im = __ldg(reinterpret_cast<const float2*>(tw + OLS_TW1_IM + 256 * k0 + c));
            v[s] = cmul(v[s], w);
        }
#else
        cp w1;
        w1.re = *reinterpret_cast<const float2*>(tw + c);
        w1.im = *reinterpret_cast<const float2*>(tw + 256 + c);
        apply_twiddles<true>(v, w1);
#endif
        const int g = t >> 3, j = 2 * (t & 7);
        const int off = 16 * g + ((j + 4 * rot_of(g)) & 15);
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int k0 = r16_k(s);
            *reinterpret_cast<float2*>(&sre[272 * k0 + off]) = v[s].re;
            *reinterpret_cast<float2*>(&sim[272 * k0 + off]) = v[s].im;
        }
    }
    __syncthreads();
    // ------------------------------------------------------------------ F2: stride 16
    const int k0_2 = t >> 3, n0_2 = 2 * (t & 7);
    int off2[4];
#pragma unroll
    for (int r = 0; r < 4; r++) off2[r] = 272 * k0_2 + ((n0_2 + 4 * r) & 15);
    const float4* tw2 = reinterpret_cast<const float4*>(tw + OLS_TW2_RE) + (n0_2 >> 1);   // tw2[8*k]: roots for (k, n0), (k, n0+1)
    {
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const int a = 16 * n1 + off2[(n1 >> 1) & 3];
            v[n1].re = *reinterpret_cast<const float2*>(&sre[a]);
            v[n1].im = *reinterpret_cast<const float2*>(&sim[a]);
        }
        r16<false>(v);
#if OLS_COMPUTE_TW2
        {
            const float4 f = __ldg(tw2 + 8);
            cp w1;
            w1.re = make_float2(f.x, f.y);
            w1.im = make_float2(f.z, f.w);
            apply_twiddles<true>(v, w1);
        }
#else
#pragma unroll
        for (int s = 1; s < 16; s++) {
            const int k1 = r16_k(s);
            const float4 f = __ldg(tw2 + 8 * k1);
            cp w;
            w.re = make_float2(f.x, f.y);
            w.im = make_float2(f.z, f.w);
            v[s] = cmul(v[s], w);
        }
#endif
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int k1 = r16_k(s);
            const int a = 16 * k1 + off2[(k1 >> 1) & 3];
            *reinterpret_cast<float2*>(&sre[a]) = v[s].re;
            *reinterpret_cast<float2*>(&sim[a]) = v[s].im;
        }
    }
#if OLS_WARP_LOCAL
    __syncwarp();
#else
    __syncthreads();
#endif
    // ------------------------------------------------------------------ F3 | *H | I3 on two groups of 16 contiguous points
#if OLS_UNROLL_F3
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int half = 0; half < 2; half++) {
#if OLS_WARP_LOCAL
        const int k0 = t >> 3, g = (t & 7) + 8 * half;
#else
        const int gg = t + 128 * half;           // group index 0..255
        const int k0 = gg >> 4, g = gg & 15;
#endif
        const int r = rot_of(g);
        const int base = 272 * k0 + 16 * g;
        cp P[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int a = base + 4 * ((q + r) & 3);
            const float4 fr = *reinterpret_cast<const float4*>(&sre[a]);
            const float4 fi = *reinterpret_cast<const float4*>(&sim[a]);
            P[2 * q].re = make_float2(fr.x, fr.y); P[2 * q + 1].re = make_float2(fr.z, fr.w);
            P[2 * q].im = make_float2(fi.x, fi.y); P[2 * q + 1].im = make_float2(fi.z, fi.w);
        }
        fft16_dif_fwd(P);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            // plan layout: [half][q][thread][4] -> a warp reads 512 contiguous bytes per request
#if OLS_H_TEX
            // texture path: keeps these 256 wavefronts per block off the LSU data pipe (the kernel's bottleneck)
            const float4 hr = tex1Dfetch<float4>(htex, (half * 4 + q) * OLS_THREADS + t);
            const float4 hi = tex1Dfetch<float4>(htex, 1024 + (half * 4 + q) * OLS_THREADS + t);
#else
            const float4 hr = __ldg(reinterpret_cast<const float4*>(Hre + ((half * 4 + q) * OLS_THREADS + t) * 4));
            const float4 hi = __ldg(reinterpret_cast<const float4*>(Him + ((half * 4 + q) * OLS_THREADS + t) * 4));
#endif
            cp h0, h1;
            h0.re = make_float2(hr.x, hr.y); h0.im = make_float2(hi.x, hi.y);
            h1.re = make_float2(hr.z, hr.w); h1.im = make_float2(hi.z, hi.w);
            P[2 * q] = cmul(P[2 * q], h0);
            P[2 * q + 1] = cmul(P[2 * q + 1], h1);
        }
        fft16_dit_inv(P);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int a = base + 4 * ((q + r) & 3);
            // 64-bit stores: a 128-bit store would need the four values in one aligned register quad (4 MOVs)
            sts64(&sre[a], P[2 * q].re);
            sts64(&sre[a + 2], P[2 * q + 1].re);
            sts64(&sim[a], P[2 * q].im);
            sts64(&sim[a + 2], P[2 * q + 1].im);
        }
    }
#if OLS_WARP_LOCAL
    __syncwarp();
#else
    __syncthreads();
#endif
    // ------------------------------------------------------------------ I2: stride 16 (DIT: twiddle first)
    {
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            const int a = 16 * k1 + off2[(k1 >> 1) & 3];
            v[k1].re = *reinterpret_cast<const float2*>(&sre[a]);
            v[k1].im = *reinterpret_cast<const float2*>(&sim[a]);
        }
#if OLS_COMPUTE_TW2
        {
            const float4 f = __ldg(tw2 + 8);
            cp w1;
            w1.re = make_float2(f.x, f.y);
            w1.im = pneg(make_float2(f.z, f.w));
            apply_twiddles<false>(v, w1);
        }
#else
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) {
            const float4 f = __ldg(tw2 + 8 * k1);
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
            const int a = 16 * n1 + off2[(n1 >> 1) & 3];
            *reinterpret_cast<float2*>(&sre[a]) = v[s].re;
            *reinterpret_cast<float2*>(&sim[a]) = v[s].im;
        }
    }
    __syncthreads();
    // ------------------------------------------------------------------ I1: stride 256, valid outputs to global
    {
        const int c = 2 * t;
        const int g = t >> 3, j = 2 * (t & 7);
        const int off = 16 * g + ((j + 4 * rot_of(g)) & 15);
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++) {
            v[k0].re = *reinterpret_cast<const float2*>(&sre[272 * k0 + off]);
            v[k0].im = *reinterpret_cast<const float2*>(&sim[272 * k0 + off]);
        }
#if OLS_TABLE_TW1
#pragma unroll
        for (int k0 = 1; k0 < 16; k0++) {
            cp w;
            w.re = __ldg(reinterpret_cast<const float2*>(tw + OLS_TW1_RE + 256 * k0 + c));
            w.im = __ldg(reinterpret_cast<const float2*>(tw + OLS_TW1_IM + 256 * k0 + c));
            v[k0] = cmul_conj(v[k0], w);
        }
#else
        cp w1;
        w1.re = *reinterpret_cast<const float2*>(tw + c);
        w1.im = pneg(*reinterpret_cast<const float2*>(tw + 256 + c));
        apply_twiddles<false>(v, w1);
#endif
        r16<true>(v);
        // output i = i0 + m, m = c + 256*n2 - m_first in [0, step) and i < N
        const int mlo = c - m_first;                       // m for n2 = 0
        int mhi = step;                                    // exclusive bound on m
        if (i0 + step > N) mhi = N - i0;                   // last block of the vector
        float2* py = yr + i0 + mlo;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const int n2 = r16_k(s);
            const int m = mlo + 256 * n2;
            if (ALIGNED) {
                if (m >= 0 && m < mhi)
                    *reinterpret_cast<float4*>(py + 256 * n2) = make_float4(v[s].re.x, v[s].im.x, v[s].re.y, v[s].im.y);
            } else {
                if (m >= 0 && m < mhi) py[256 * n2] = make_float2(v[s].re.x, v[s].im.x);
                if (m + 1 >= 0 && m + 1 < mhi) py[256 * n2 + 1] = make_float2(v[s].re.y, v[s].im.y);
            }
        }
    }
}

// Hpos (planar, kernel layout) <- Hs (interleaved, natural order, already scaled by 1/M), delayed by d
// samples: H_d[k] = H[k] * exp(-2 pi i k d / M).  Kernel layout: value of block position
// p = 16*gg + 4*q + e (gg = 16*k0 + g, owned by thread t = 8*k0 + (g & 7) in iteration half = g >> 3) is stored at
// ((half*4 + q)*128 + t)*4 + e.
__global__ void ols4096_permute_h_kernel(const float2* __restrict__ Hs, float* __restrict__ Hre, float* __restrict__ Him, int d) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= OLS_M) return;
    const int k = freq_of_pos(p);
    float2 h = Hs[k];
    if (d) h = cmul(h, unit_root<float>((unsigned long long)k * (unsigned long long)d, OLS_M, -1));
    const int gg = p >> 4, q = (p >> 2) & 3, e = p & 3;
#if OLS_WARP_LOCAL
    const int t = 8 * (gg >> 4) + (gg & 7), half = (gg >> 3) & 1;
#else
    const int t = gg & 127, half = gg >> 7;
#endif
    const int dst = ((half * 4 + q) * OLS_THREADS + t) * 4 + e;
    Hre[dst] = h.x;
    Him[dst] = h.y;
}

namespace {
std::mutex g_tw_mu;
std::map<int, float*> g_tw;  // per device
}

static const float* ols4096_twiddles() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { set_last_error("ols4096: cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_tw_mu);
    auto it = g_tw.find(d);
    if (it != g_tw.end()) return it->second;
    std::vector<float> h(OLS_TW_FLOATS);
    for (int c = 0; c < 256; c++) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)c / 4096.0L;
        h[c] = (float)cosl(a);
        h[256 + c] = (float)sinl(a);
    }
    for (int k = 0; k < 16; k++)
        for (int n = 0; n < 16; n++) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)((k * n) % 256) / 256.0L;
            h[OLS_TW2_RE + (8 * k + n / 2) * 4 + (n & 1)] = (float)cosl(a);
            h[OLS_TW2_RE + (8 * k + n / 2) * 4 + 2 + (n & 1)] = (float)sinl(a);
        }
    for (int k = 0; k < 16; k++)
        for (int c = 0; c < 256; c++) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)((k * c) % 4096) / 4096.0L;
            h[OLS_TW1_RE + 256 * k + c] = (float)cosl(a);
            h[OLS_TW1_IM + 256 * k + c] = (float)sinl(a);
        }
    float* dev = nullptr;
    if (cudaMalloc(&dev, OLS_TW_FLOATS * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(dev, h.data(), OLS_TW_FLOATS * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        if (dev) cudaFree(dev);
        set_last_error("ols4096: twiddle table allocation failed");
        return nullptr;
    }
    g_tw[d] = dev;
    return dev;
}

bool ols4096_applicable(size_t N, size_t L, size_t M) {
    // needs one wrap at most per strided load and 32-bit row indices
    return M == OLS_M && L >= 2 && L <= OLS_M / 2 - 2 && N >= OLS_M && N < (1ull << 30);
}

// plan geometry shared by prepare and convolve: delay d makes the input->output index distance even
static inline void ols4096_geometry(size_t L, int* d, int* shift, int* m_first, int* step) {
    const int cl = (int)(L - L / 2);
    *d = (cl - 1) & 1;
    *shift = cl - 1 + *d;
    int mf = (int)L - 1 + *d;      // first block position whose circular convolution value is valid
    if (mf & 1) mf++;
    *m_first = mf;
    *step = (OLS_M - mf) & ~1;
}

// Hpos: 2*4096 floats (re plane, im plane) <- Hs from ols_prepare<float>()
int ols4096_prepare(const void* Hs, void* Hpos, size_t L, cudaStream_t st) {
    int d, shift, m_first, step;
    ols4096_geometry(L, &d, &shift, &m_first, &step);
    float* hp = reinterpret_cast<float*>(Hpos);
    ols4096_permute_h_kernel<<<OLS_M / 256, 256, 0, st>>>(reinterpret_cast<const float2*>(Hs), hp, hp + OLS_M, d);
    BDSP_LAUNCHED();
    return 0;
}

int ols4096_convolve(const void* x, void* y, size_t N, size_t batch, size_t L, const void* Hpos, cudaTextureObject_t htex, cudaStream_t st) {
    if (x == y) { set_last_error("ols4096_convolve: in-place operation is not supported"); return -3; }
    int d, shift, m_first, step;
    ols4096_geometry(L, &d, &shift, &m_first, &step);
    const long long bpv = ((long long)N + step - 1) / step;
    const long long grid = bpv * (long long)batch;
    if (grid > 0x7fffffffll) { set_last_error("ols4096_convolve: grid too large"); return -2; }
    const float* hp = reinterpret_cast<const float*>(Hpos);
    const float* tw = ols4096_twiddles();
    if (!tw) return -1;
    const cudaTextureObject_t xtex = 0;
    const bool aligned = (N % 2 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    if (aligned)
        ols4096_kernel<true><<<(unsigned)grid, OLS_THREADS, 0, st>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(y), (int)N,
                                                                     m_first, step, shift, (int)bpv, hp, hp + OLS_M, tw, htex, xtex);
    else
        ols4096_kernel<false><<<(unsigned)grid, OLS_THREADS, 0, st>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(y), (int)N,
                                                                      m_first, step, shift, (int)bpv, hp, hp + OLS_M, tw, htex, xtex);
    BDSP_LAUNCHED();
    return 0;
}

}  // namespace bdsp
