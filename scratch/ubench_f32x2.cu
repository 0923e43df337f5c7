// Microbenchmark: scalar FFMA vs packed FFMA2/FADD2 throughput on sm_100a, and co-issue with LDS.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template<int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    __shared__ float2 sm[256*4];
    float2 acc[8];
    #pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = make_float2(threadIdx.x + i, i);
    float2 A = make_float2(a, a * 0.5f), B = make_float2(b, b * 2.f);
    sm[threadIdx.x] = A; sm[threadIdx.x+256]=B; sm[threadIdx.x+512]=A; sm[threadIdx.x+768]=B;
    __syncthreads();
    float2 ld = make_float2(0,0);
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); }
            if (MODE == 1) { acc[i] = __ffma2_rn(acc[i], A, B); }
            if (MODE == 2) { acc[i].x = acc[i].x + A.x; acc[i].y = acc[i].y + A.y; }
            if (MODE == 3) { acc[i] = __fadd2_rn(acc[i], A); }
            if (MODE == 4) { acc[i] = __ffma2_rn(acc[i], A, B); if ((i & 1) == 0) { float2 t = sm[(threadIdx.x + it*8 + i*32) & 1023]; ld.x += t.x; ld.y += t.y; } }
            if (MODE == 5) { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); if ((i & 1) == 0) { float2 t = sm[(threadIdx.x + it*8 + i*32) & 1023]; ld.x += t.x; ld.y += t.y; } }
            if (MODE == 6) { acc[i] = __fmul2_rn(acc[i], A); }
        }
    }
    float s = ld.x + ld.y;
    #pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int MODE> void run(const char* name, float* d) {
    int blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double lane_ops = (double)blocks * 256 * ITERS * 16;  // 16 scalar lane-ops per iter (8 float2)
    printf("%-28s %8.3f ms  %8.2f Tlane-op/s (x2 flops for fma)\n", name, ms, lane_ops / ms / 1e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA", d); run<1>("packed FFMA2", d); run<2>("scalar FADD", d); run<3>("packed FADD2", d);
    run<6>("packed FMUL2", d);
    run<5>("scalar FFMA + LDS.64", d); run<4>("packed FFMA2 + LDS.64", d);
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
