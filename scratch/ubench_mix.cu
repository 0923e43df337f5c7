// Can scalar FP32 ops (fmalite pipe) run concurrently with packed FP32x2 ops (fmaheavy pipe)?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
// NP packed + NS scalar independent accumulators per iteration
template<int NP, int NS>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float2 pa[NP > 0 ? NP : 1]; float sa[NS > 0 ? NS : 1];
    for (int i = 0; i < NP; i++) pa[i] = make_float2(threadIdx.x + i, i);
    for (int i = 0; i < NS; i++) sa[i] = threadIdx.x * 0.5f + i;
    float2 A = make_float2(a, a * 0.5f), B = make_float2(b, b * 2.f);
    for (int it = 0; it < ITERS; it++) {
        #pragma unroll
        for (int i = 0; i < (NP > NS ? NP : NS); i++) {
            if (i < NP) pa[i] = __ffma2_rn(pa[i], A, B);
            if (i < NS) sa[i] = fmaf(sa[i], a, b);
        }
    }
    float s = 0;
    for (int i = 0; i < NP; i++) s += pa[i].x + pa[i].y;
    for (int i = 0; i < NS; i++) s += sa[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int NP, int NS> void run(float* d) {
    int blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NP,NS><<<blocks, 256>>>(d, 1.0001f, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<NP,NS><<<blocks, 256>>>(d, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double lane_ops = (double)blocks * 256 * ITERS * (2.0 * NP + NS);
    double instr = (double)blocks * 8 * ITERS * (NP + NS);  // warp instructions
    printf("packed %2d scalar %2d: %7.3f ms  %6.2f Tlane-fma/s  %5.2f warp-instr/clk/SMSP(@1.965GHz)\n", NP, NS, ms, lane_ops / ms / 1e9, instr / (ms * 1e-3) / (148.0 * 4 * 1.965e9));
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<8,0>(d); run<0,8>(d); run<0,16>(d); run<8,4>(d); run<8,8>(d); run<6,6>(d); run<8,2>(d); run<4,8>(d); run<12,4>(d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
