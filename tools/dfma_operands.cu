// dfma_operands.cu -- does the DFMA issue rate on B200 depend on how many DISTINCT register operands an instruction reads?
// (k_score's hot loop sits at 66 % of the fp64 pipe with math_pipe_throttle as its top stall: profiles/r02_*.)
// Variants, 8 independent chains per thread, 512 threads per CTA, 1 CTA per SM:
//   r1   acc = fma(acc, a, b)        one varying register pair (what tools/fp64_peak.cu measures)
//   r2   acc = fma(acc, x_i, b)      two
//   r3   acc = fma(acc, x_i, y_i)    three distinct register pairs per instruction
//   r3s  acc = fma(x_i, y_i, acc) with x_i, y_i swapped between slots every other instruction
//   ur   acc = fma(acc, x_i, <uniform/constant operand>)
//   mix  r3 interleaved with one integer instruction per DFMA (the loop issues 37 other instructions per 44 fp64)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tools/dfma_operands tools/dfma_operands.cu
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
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 1) acc[i] = fma(acc[i], a, b);
            if (MODE == 2) acc[i] = fma(acc[i], x[i], b);
            if (MODE == 3) acc[i] = fma(acc[i], x[i], y[i]);
            if (MODE == 4) acc[i] = (i & 1) ? fma(x[i], y[i], acc[i]) : fma(y[i], x[i], acc[i]);
            if (MODE == 5) acc[i] = fma(acc[i], x[i], 1.0000001);
            if (MODE == 6) { acc[i] = fma(acc[i], x[i], y[i]); cnt[i] += unsigned(__double2hiint(acc[i])) >> 31; }
            if (MODE == 7) { acc[i] = fma(acc[i], a, b); cnt[i] += unsigned(__double2hiint(acc[i])) >> 31; }
            if (MODE == 8) acc[i] = fma(x[i], y[(i + 1) & 7], acc[i]);       // accumulate form: two operands shared with neighbours
            if (MODE == 9) acc[i] = fma(x[i & 1], y[i], acc[i]);             // one operand repeats in the same slot (reuse cache candidate)
        }
    }
    double s = 0;
    unsigned c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += acc[i]; c += cnt[i]; }
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
    const double flop = double(sms) * 512 * iters * 8 * 2;
    const char* names[] = {"", "r1 fma(acc,a,b)", "r2 fma(acc,x_i,b)", "r3 fma(acc,x_i,y_i)", "r3s slots swapped", "ur fma(acc,x_i,imm)",
                           "r3 + int op", "r1 + int op", "fma(x_i,y_j,acc)", "fma(x_{i&1},y_i,acc)"};
    float ms;
#define RUN(M) ms = run<M>(out, in, sms, iters); printf("%-24s %8.3f ms  %6.2f TFLOP/s  %5.1f %% of 37.2\n", names[M], ms, flop / ms / 1e9, flop / ms / 1e9 / 37.2 * 100);
    RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    CK(cudaDeviceSynchronize());
    return 0;
}
