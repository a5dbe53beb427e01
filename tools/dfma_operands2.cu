// dfma_operands2.cu -- second round: operand reuse cache, integer instructions beside DFMA, DADD with uniform operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tools/dfma_operands2 tools/dfma_operands2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, const double* in, int iters, double a, double b) {
    double acc[8], x[8], y[8];
    unsigned cnt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = threadIdx.x * 1e-3 + i;
        x[i] = in[(threadIdx.x + i) & 63];
        y[i] = in[(threadIdx.x + 8 + i) & 63];
        cnt[i] = threadIdx.x + i;
    }
    const double x0 = x[0], y0 = y[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 1) acc[i] = fma(x0, y[i], acc[i]);                 // slot A the same register in all 8: reuse cache?
            if (MODE == 2) acc[i] = fma(x0, y0, acc[i]);                   // slots A and B repeat
            if (MODE == 3) { acc[i] = fma(x0, y[i], acc[i]); cnt[i] = cnt[i] * 3 + 1; }   // reuse candidate + an integer op in between
            if (MODE == 4) { acc[i] = fma(acc[i], a, b); cnt[i] = cnt[i] * 3 + 1; }       // 1 register operand + IMAD (1 reg)
            if (MODE == 5) { acc[i] = fma(acc[i], a, b); cnt[i] = (cnt[i] ^ cnt[(i + 1) & 7]) + cnt[(i + 3) & 7]; }  // + LOP3/IADD3 3 regs
            if (MODE == 6) acc[i] = acc[i] + y[i];                         // DADD 2 regs
            if (MODE == 7) acc[i] = fma(acc[i], acc[i], y[i]);             // same register twice + one
            if (MODE == 8) { acc[i] = fma(acc[i], a, b); asm volatile("" ::: "memory"); }
            if (MODE == 9) { acc[i] = fma(acc[i], x[i], y[i]); }
        }
        if (MODE == 9) {   // r3 chains, but y rotates so that nothing repeats in a slot
            double t = y[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) y[i] = y[i + 1];
            y[7] = t;
        }
    }
    double s = 0;
    unsigned c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += acc[i] + y[i]; c += cnt[i]; }
    if (s == 12345.678 || c == 0x7fffffffu) out[0] = s + c;
}

template <int MODE>
float run(double* out, const double* in, int grid, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 512>>>(out, in, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 512>>>(out, in, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
    double* in; CK(cudaMalloc(&in, sizeof(h))); CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    const int iters = 20000;
    const char* names[] = {"", "fma(x0,y_i,acc_i)", "fma(x0,y0,acc_i)", "fma(x0,y_i,acc_i)+IMAD", "fma(acc,a,b)+IMAD", "fma(acc,a,b)+LOP3/IADD3",
                           "DADD acc+y_i", "fma(acc,acc,y_i)", "fma(acc,a,b) fenced", "r3"};
    float ms;
    // clk per (DFMA [+ int op]) per SM sub-partition: 16 warps / 4 = 4 warps each, 8 * iters instruction groups per warp
    const double clk_per_ms = p.clockRate;   // kHz = clk per ms
#define RUN(M) ms = run<M>(out, in, sms, iters); printf("%-26s %8.3f ms  %5.2f clk per group per sub-partition\n", names[M], ms, ms * clk_per_ms / (4.0 * 8 * iters));
    RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    CK(cudaDeviceSynchronize());
    return 0;
}
