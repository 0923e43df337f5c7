// Packed-FP32x2 complex arithmetic in INTERLEAVED form: one 64-bit register pair holds (re, im) of ONE point.
//
// Blackwell's FADD2 / FMUL2 / FFMA2 take, per operand, a lane swap (.LO_HI), per-lane signs (.NP / .PN), a 32-bit
// broadcast (.F32) or a 32-bit immediate.  With those modifiers every complex operation of an FFT is as cheap on
// interleaved data as on planar data,
//     a +- b            1 FADD2          a -+ i b          1 FADD2 (swap + lane signs on b)
//     a * w             1 FMUL2 + 1 FFMA2 (w.re / w.im broadcast from a register, a uniform register or an immediate)
// and nothing has to be re-paired: a 128-bit global or shared access moves two whole points, internal twiddles are
// immediates, and the last butterfly level of a transform over contiguous points is an ordinary packed one.  Round 1's
// planar kernels (ols4096.cuh) paid for that with ~8 MOVs per 128-bit load, UMOV pairs for every constant and scalar
// FADDs for the last level (10-15 % of all issued instructions; the kernels are issue/latency bound).
#pragma once
#include "common.cuh"

namespace bdsp {
namespace cx {

typedef float2 c2;

__device__ __forceinline__ c2 add(c2 a, c2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ c2 sub(c2 a, c2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a - i b
__device__ __forceinline__ c2 add_mi(c2 a, c2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
// a + i b
__device__ __forceinline__ c2 add_pi(c2 a, c2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
// a * (wr + i wi)
__device__ __forceinline__ c2 mulc(c2 a, float wr, float wi) {
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(wi, wi), __fmul2_rn(a, make_float2(wr, wr)));
}
__device__ __forceinline__ c2 mul(c2 a, c2 w) { return mulc(a, w.x, w.y); }
// a * conj(w)
__device__ __forceinline__ c2 mul_conj(c2 a, c2 w) { return mulc(a, w.x, -w.y); }
__device__ __forceinline__ c2 conj(c2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ c2 scale(c2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }

// radix-4 butterfly, results in natural order (a_k = sum_j a_j exp(-+2 pi i j k / 4))
template <bool INV> __device__ __forceinline__ void r4(c2& a0, c2& a1, c2& a2, c2& a3) {
    const c2 s0 = add(a0, a2), d0 = sub(a0, a2), s1 = add(a1, a3), d1 = sub(a1, a3);
    a0 = add(s0, s1);
    a2 = sub(s0, s1);
    if (!INV) { a1 = add_mi(d0, d1); a3 = add_pi(d0, d1); }
    else      { a1 = add_pi(d0, d1); a3 = add_mi(d0, d1); }
}

#define CX_C1 0.92387953251128675613f
#define CX_S1 0.38268343236508977173f
#define CX_H 0.70710678118654752440f

// v * exp(-+2 pi i m / 16); the constants become 32-bit immediates of the packed instructions
template <int MM, bool INV> __device__ __forceinline__ c2 mul_w16(c2 v) {
    constexpr int m = MM & 15;
    if constexpr (m == 0) return v;
    else if constexpr (m == 4) return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
    else if constexpr (m == 8) return make_float2(-v.x, -v.y);
    else if constexpr (m == 12) return INV ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
    else {
        constexpr float cs[16] = {1.f, CX_C1, CX_H, CX_S1, 0.f, -CX_S1, -CX_H, -CX_C1, -1.f, -CX_C1, -CX_H, -CX_S1, 0.f, CX_S1, CX_H, CX_C1};
        constexpr float sn[16] = {0.f, CX_S1, CX_H, CX_C1, 1.f, CX_C1, CX_H, CX_S1, 0.f, -CX_S1, -CX_H, -CX_C1, -1.f, -CX_C1, -CX_H, -CX_S1};
        return mulc(v, cs[m], INV ? sn[m] : -sn[m]);
    }
}

// cos(2 pi m / 64), m = 0..16
__device__ __forceinline__ constexpr float cos64_q(int m) {
    constexpr float t[17] = {1.00000000000000000000f, 0.99518472667219692873f, 0.98078528040323043058f, 0.95694033573220882438f,
                             0.92387953251128673848f, 0.88192126434835504956f, 0.83146961230254523567f, 0.77301045336273699338f,
                             0.70710678118654757274f, 0.63439328416364548779f, 0.55557023301960228867f, 0.47139673682599780857f,
                             0.38268343236508983729f, 0.29028467725446233105f, 0.19509032201612833135f, 0.09801714032956077016f, 0.f};
    return t[m];
}
__device__ __forceinline__ constexpr float cos64(int m) {   // any m
    m &= 63;
    return m <= 16 ? cos64_q(m) : m <= 32 ? -cos64_q(32 - m) : m <= 48 ? -cos64_q(m - 32) : cos64_q(64 - m);
}
__device__ __forceinline__ constexpr float sin64(int m) { return cos64(m - 16); }
// v * exp(-+2 pi i m / 64)
template <int MM, bool INV> __device__ __forceinline__ c2 mul_w64(c2 v) {
    constexpr int m = MM & 63;
    if constexpr (m % 16 == 0) return mul_w16<m / 4, INV>(v);
    else return mulc(v, cos64(m), INV ? sin64(m) : -sin64(m));
}
template <int LO, int HI, bool INV> struct W64Sel {   // run-time (but compile-time constant after unrolling) index -> template
    static __device__ __forceinline__ c2 apply(int m, c2 v) {
        if constexpr (LO == HI) return mul_w64<LO, INV>(v);
        else {
            constexpr int MID = (LO + HI) / 2;
            return m <= MID ? W64Sel<LO, MID, INV>::apply(m, v) : W64Sel<MID + 1, HI, INV>::apply(m, v);
        }
    }
};

// 16-point transform of v[j] (j natural) over a strided dimension.  On return slot s holds frequency r16_k(s).
__device__ __forceinline__ constexpr int r16_k(int s) { return (s >> 2) + 4 * (s & 3); }

template <bool INV> __device__ __forceinline__ void r16(c2 (&v)[16]) {
#pragma unroll
    for (int j0 = 0; j0 < 4; j0++) r4<INV>(v[j0], v[4 + j0], v[8 + j0], v[12 + j0]);
    // twiddle W16^{j0*k1} on slot 4*k1 + j0
    v[5] = mul_w16<1, INV>(v[5]);   v[6] = mul_w16<2, INV>(v[6]);   v[7] = mul_w16<3, INV>(v[7]);
    v[9] = mul_w16<2, INV>(v[9]);   v[10] = mul_w16<4, INV>(v[10]); v[11] = mul_w16<6, INV>(v[11]);
    v[13] = mul_w16<3, INV>(v[13]); v[14] = mul_w16<6, INV>(v[14]); v[15] = mul_w16<9, INV>(v[15]);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) r4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// multiply v[slot] by w1^{k(slot)}, k(slot) = SLOTMAP ? r16_k(slot) : slot  (powers as A_a * B_b, k = a + 4 b)
template <bool SLOTMAP> __device__ __forceinline__ void apply_twiddles(c2 (&v)[16], c2 w1) {
    c2 A[4], B[4];
    A[1] = w1;
    A[2] = mul(w1, w1);
    A[3] = mul(A[2], w1);
    B[1] = mul(A[2], A[2]);
    B[2] = mul(B[1], B[1]);
    B[3] = mul(B[2], B[1]);
#pragma unroll
    for (int s = 1; s < 16; s++) {
        const int k = SLOTMAP ? r16_k(s) : s;
        const int a = k & 3, b = k >> 2;
        c2 w;
        if (b == 0) w = A[a];
        else if (a == 0) w = B[b];
        else w = mul(A[a], B[b]);
        v[s] = mul(v[s], w);
    }
}

__host__ __device__ __forceinline__ constexpr int bitrev4(int j) { return ((j & 1) << 3) | ((j & 2) << 1) | ((j & 4) >> 1) | ((j & 8) >> 3); }
__host__ __device__ __forceinline__ constexpr int bitrev5(int j) {
    return ((j & 1) << 4) | ((j & 2) << 2) | (j & 4) | ((j & 8) >> 2) | ((j & 16) >> 4);
}

__host__ __device__ __forceinline__ constexpr int bitrev6(int j) {
    return ((j & 1) << 5) | ((j & 2) << 3) | ((j & 4) << 1) | ((j & 8) >> 1) | ((j & 16) >> 3) | ((j & 32) >> 5);
}
template <int LOG> __host__ __device__ __forceinline__ constexpr int bitrev(int j) {
    int r = 0;
    for (int i = 0; i < LOG; i++) r |= ((j >> i) & 1) << (LOG - 1 - i);
    return r;
}

// Transform over N = 2^LOG <= 64 points held in registers.  DIF: natural in -> bit-reversed out (slot j holds
// frequency bitrev(j)); every level is a packed butterfly, the twiddles W_{2 span}^{j} = W_64^{j * 32 / span} are immediates.
template <int LOG, int LV, bool INV> struct FftLevel {   // one butterfly level with compile-time span (full unrolling)
    static constexpr int NN = 1 << LOG;
    static constexpr int SPAN = NN >> (LV + 1);
    static __device__ __forceinline__ void dif(c2 (&P)[1 << LOG]) {
#pragma unroll
        for (int b = 0; b < NN; b += 2 * SPAN) {
#pragma unroll
            for (int j = 0; j < SPAN; j++) {
                const c2 u = add(P[b + j], P[b + j + SPAN]);
                const c2 d = sub(P[b + j], P[b + j + SPAN]);
                P[b + j] = u;
                P[b + j + SPAN] = W64Sel<0, 31, INV>::apply(j * (32 / SPAN), d);
            }
        }
        if constexpr (LV + 1 < LOG) FftLevel<LOG, LV + 1, INV>::dif(P);
    }
    static __device__ __forceinline__ void dit(c2 (&P)[1 << LOG]) {
        if constexpr (LV + 1 < LOG) FftLevel<LOG, LV + 1, INV>::dit(P);
#pragma unroll
        for (int b = 0; b < NN; b += 2 * SPAN) {
#pragma unroll
            for (int j = 0; j < SPAN; j++) {
                const c2 t = W64Sel<0, 31, INV>::apply(j * (32 / SPAN), P[b + j + SPAN]);
                const c2 u = P[b + j];
                P[b + j] = add(u, t);
                P[b + j + SPAN] = sub(u, t);
            }
        }
    }
};
template <int LOG, bool INV> __device__ __forceinline__ void fft_dif(c2 (&P)[1 << LOG]) { FftLevel<LOG, 0, INV>::dif(P); }

// inverse direction of the same structure: DIT, bit-reversed in -> natural out, twiddle (conjugated roots for INV = true)
// BEFORE the butterfly
template <int LOG, bool INV> __device__ __forceinline__ void fft_dit(c2 (&P)[1 << LOG]) { FftLevel<LOG, 0, INV>::dit(P); }

// v[bitrev6(k)] *= w1^k for k = 1..63 (the slot order fft_dif<6> leaves / fft_dit<6> expects), powers as A_a B_b C_c,
// k = a + 4 b + 16 c: 62 products for the powers, 63 for the application
__device__ __forceinline__ void apply_twiddles64(c2 (&v)[64], c2 w1) {
    c2 A[4], B[4], C[4];
    A[1] = w1; A[2] = mul(w1, w1); A[3] = mul(A[2], w1);
    B[1] = mul(A[2], A[2]); B[2] = mul(B[1], B[1]); B[3] = mul(B[2], B[1]);
    C[1] = mul(B[2], B[2]); C[2] = mul(C[1], C[1]); C[3] = mul(C[2], C[1]);
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            c2 ab;
            if (a == 0 && b == 0) ab = make_float2(1.f, 0.f);
            else if (b == 0) ab = A[a];
            else if (a == 0) ab = B[b];
            else ab = mul(A[a], B[b]);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int k = a + 4 * b + 16 * c;
                if (k == 0) continue;
                c2 w;
                if (a == 0 && b == 0) w = C[c];
                else if (c == 0) w = ab;
                else w = mul(ab, C[c]);
                v[bitrev6(k)] = mul(v[bitrev6(k)], w);
            }
        }
}

}  // namespace cx
}  // namespace bdsp
