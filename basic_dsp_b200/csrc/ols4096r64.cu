// Fused 4096-point overlap-save block for the fast convolution hot path (the reference's overlap_discard,
// convolution.rs:304-461: 2 rustfft calls + a scalar multiply loop per block, the spectrum going through memory):
// forward FFT, spectrum multiply and inverse FFT of one block in ONE kernel with the block resident in shared memory.
//
// This version is built around what bounds the kernel on B200.  Ablation of the radix-16 version (ols4096i.cu,
// profiles/r2_ols_ablation.txt): with ALL arithmetic removed the block still takes 79 % of its time - the kernel is bound
// by the L1/shared-memory data pipe (1 wavefront of 128 bytes per clock and SM), to which every 128-bit shared access
// contributes 4 wavefronts per warp, and FP32 work only shows through where it fails to overlap.  A 4096-point block as
// 16 x 16 x 16 needs FOUR exchanges through shared memory (2048 wavefronts per block, 60 % of the pipe's load).
// As 64 x 64 it needs TWO:
//   F1  radix-64 DIF down the 64 columns (one column per thread, inputs straight from global memory)   -> smem | CTA barrier
//   F2  radix-64 DIF along the 64 contiguous points of a row | * H | radix-64 DIT   (all in registers)     -> smem | CTA barrier
//   I1  radix-64 DIT down the columns, valid outputs straight to global memory
// 64 threads per block, 64 points per thread (128 data registers), interleaved-complex FP32x2 arithmetic (cxmath.cuh)
// with all internal twiddles as immediates; one set of inter-stage twiddles per direction instead of two.
// Shared layout: point (row r, column c) lives in float2 slot 64 r + (c ^ 2 (r & 7)): the 16-byte chunks of a row are
// XOR-swizzled by the row number, so the 128-bit row accesses of a quarter-warp (8 consecutive rows) fall into 8 different
// bank windows and the 64-bit column accesses of a warp cover whole 128-byte lines.  No padding: 32 KB per CTA.
// The plan stores H in the order the forward transform leaves the spectrum in (position order): no reordering pass.
#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "conv.cuh"
#include "cxmath.cuh"

namespace bdsp {

using namespace cx;

#define O6_M 4096
#define O6_T 64
#ifndef O6_MIN_CTAS
#define O6_MIN_CTAS 5
#endif

__global__ void __launch_bounds__(O6_T, O6_MIN_CTAS)
ols4096r64_kernel(const float2* __restrict__ x, float2* __restrict__ y, int N, int m_first, int step, int shift,
                  int blocks_per_vec, const float2* __restrict__ tw, cudaTextureObject_t htex) {
    __shared__ __align__(16) float2 sm[O6_M];
    const int t = threadIdx.x;
    const int vec = blockIdx.x / blocks_per_vec;
    const int blk = blockIdx.x - vec * blocks_per_vec;
    const int i0 = blk * step;
    const float2* xr = x + (size_t)vec * (size_t)N;
    float2* yr = y + (size_t)vec * (size_t)N;

    c2 v[64];
    // ------------------------------------------------------------------ F1: down column t (block position p = 64 a + t)
    {
        // block position p holds x[(p0 + p) mod N], p0 = i0 + shift - m_first  (p0 > -4096)
        const int p0 = i0 + shift - m_first;
        if (p0 >= 0 && p0 + O6_M <= N) {   // block-uniform: no wrap-around inside this block
            const float2* px = xr + p0 + t;
#pragma unroll
            for (int a = 0; a < 64; a++) v[a] = __ldg(px + 64 * a);
        } else {                             // first / last block of a vector: circular indexing
            int idx = (p0 + t) % N;
            if (idx < 0) idx += N;
            const int adv = 64 % N;
#pragma unroll
            for (int a = 0; a < 64; a++) {
                v[a] = __ldg(&xr[idx]);
                idx += adv; if (idx >= N) idx -= N;
            }
        }
        fft_dif<6, false>(v);                // slot s holds k0 = bitrev6(s)
        apply_twiddles64(v, __ldg(tw + t));  // * W4096^{k0 t}
#pragma unroll
        for (int s = 0; s < 64; s++) {
            const int k0 = bitrev6(s);
            sm[64 * k0 + (t ^ (2 * (k0 & 7)))] = v[s];
        }
    }
    __syncthreads();
    // ------------------------------------------------------------------ F2 | *H | I2 along row t (frequency k = t + 64 k1)
    {
        float2* row = sm + 64 * t;
        const int sw = t & 7;
#pragma unroll
        for (int q = 0; q < 32; q++) {
            const float4 f = *reinterpret_cast<const float4*>(row + 2 * (q ^ sw));
            v[2 * q] = make_float2(f.x, f.y);
            v[2 * q + 1] = make_float2(f.z, f.w);
        }
        fft_dif<6, false>(v);                // slot j holds k1 = bitrev6(j)
#pragma unroll
        for (int q = 0; q < 32; q++) {
            // plan layout: [q][thread] float4 = H of row t, slots 2q and 2q + 1 -> a warp reads 512 contiguous bytes
            const float4 h = tex1Dfetch<float4>(htex, q * O6_T + t);
            v[2 * q] = mul(v[2 * q], make_float2(h.x, h.y));
            v[2 * q + 1] = mul(v[2 * q + 1], make_float2(h.z, h.w));
        }
        fft_dit<6, true>(v);
#pragma unroll
        for (int q = 0; q < 32; q++)
            *reinterpret_cast<float4*>(row + 2 * (q ^ sw)) = make_float4(v[2 * q].x, v[2 * q].y, v[2 * q + 1].x, v[2 * q + 1].y);
    }
    __syncthreads();
    // ------------------------------------------------------------------ I1: down column t, valid outputs to global
    {
#pragma unroll
        for (int s = 0; s < 64; s++) {
            const int k0 = bitrev6(s);
            v[s] = sm[64 * k0 + (t ^ (2 * (k0 & 7)))];
        }
        apply_twiddles64(v, conj(__ldg(tw + t)));
        fft_dit<6, true>(v);                 // natural order: v[a] is block position 64 a + t
        // output i = i0 + m, m = t + 64 a - m_first in [0, step) and i < N
        const int mlo = t - m_first;                       // m for a = 0
        int mhi = step;                                    // exclusive bound on m
        if (i0 + step > N) mhi = N - i0;                   // last block of the vector
        float2* py = yr + i0 + mlo;
#pragma unroll
        for (int a = 0; a < 64; a++) {
            const int m = mlo + 64 * a;
            if (m >= 0 && m < mhi) py[64 * a] = v[a];
        }
    }
}

// frequency index held at block position p = 64 r + j after F1 and F2
__host__ __device__ __forceinline__ int o6_freq_of_pos(int p) { return (p >> 6) + 64 * bitrev6(p & 63); }

// Hpos (kernel layout) <- Hs (interleaved, natural order, already scaled by 1/M).  Position p = 64 r + j (row r = thread,
// slot j) is stored at float2 index 2 * ((j / 2) * 64 + r) + (j & 1).
__global__ void ols4096r64_permute_h_kernel(const float2* __restrict__ Hs, float2* __restrict__ Hpos) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= O6_M) return;
    const int r = p >> 6, j = p & 63;
    Hpos[2 * ((j >> 1) * O6_T + r) + (j & 1)] = Hs[o6_freq_of_pos(p)];
}

namespace {
std::mutex g_o6_mu;
std::map<int, float2*> g_o6_tw;  // per device: W4096^c, c in [0, 64)

const float2* ols4096r64_twiddles() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { set_last_error("ols4096: cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_o6_mu);
    auto it = g_o6_tw.find(d);
    if (it != g_o6_tw.end()) return it->second;
    std::vector<float2> h(64);
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    for (int c = 0; c < 64; c++) {
        h[c].x = (float)cosl(-tau * (long double)c / 4096.0L);
        h[c].y = (float)sinl(-tau * (long double)c / 4096.0L);
    }
    float2* dev = nullptr;
    if (cudaMalloc(&dev, 64 * sizeof(float2)) != cudaSuccess ||
        cudaMemcpy(dev, h.data(), 64 * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) {
        if (dev) cudaFree(dev);
        cudaGetLastError();
        set_last_error("ols4096: twiddle table allocation failed");
        return nullptr;
    }
    g_o6_tw[d] = dev;
    return dev;
}
}  // namespace

bool ols4096_applicable(size_t N, size_t L, size_t M) {
    // needs one wrap at most per strided load and 32-bit row indices
    return M == O6_M && L >= 2 && L <= O6_M / 2 && N >= O6_M && N < (1ull << 30);
}

// plan geometry: block position m holds the circular convolution value of output i0 + m - m_first for m >= m_first = L - 1
// (every access moves whole points, so no alignment delay is needed)
static inline void ols4096_geometry(size_t L, int* shift, int* m_first, int* step) {
    const int cl = (int)(L - L / 2);
    *shift = cl - 1;
    *m_first = (int)L - 1;
    *step = O6_M - *m_first;
}

// Hpos: 4096 float2 <- Hs from the plan (FFT_4096(pad(h)) / 4096, natural order)
int ols4096_prepare(const void* Hs, void* Hpos, size_t L, cudaStream_t st) {
    (void)L;
    ols4096r64_permute_h_kernel<<<O6_M / 256, 256, 0, st>>>(reinterpret_cast<const float2*>(Hs), reinterpret_cast<float2*>(Hpos));
    BDSP_LAUNCHED();
    return 0;
}

int ols4096_convolve(const void* x, void* y, size_t N, size_t batch, size_t L, const void* Hpos, cudaTextureObject_t htex, cudaStream_t st) {
    (void)Hpos;
    if (x == y) { set_last_error("ols4096_convolve: in-place operation is not supported"); return -3; }
    int shift, m_first, step;
    ols4096_geometry(L, &shift, &m_first, &step);
    const long long bpv = ((long long)N + step - 1) / step;
    const long long grid = bpv * (long long)batch;
    if (grid > 0x7fffffffll) { set_last_error("ols4096_convolve: grid too large"); return -2; }
    const float2* tw = ols4096r64_twiddles();
    if (!tw) return -1;
    ols4096r64_kernel<<<(unsigned)grid, O6_T, 0, st>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(y), (int)N,
                                                       m_first, step, shift, (int)bpv, tw, htex);
    BDSP_LAUNCHED();
    return 0;
}

}  // namespace bdsp
