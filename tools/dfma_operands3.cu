// dfma_operands3.cu -- marginal cost of ONE non-fp64 instruction beside a saturated DFMA stream, per instruction kind:
// which of the integer / load / MUFU instructions of k_score's hot loop are free and which take time from the fp64 pipe's issue.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tools/dfma_operands3 tools/dfma_operands3.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, const double* in, int iters, double a, double b) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    double acc[8];
    unsigned cnt[8], aux[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x * 1e-3 + i; cnt[i] = threadIdx.x + i; aux[i] = threadIdx.x * 7 + i; }
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] = fma(acc[i], a, b);
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, 3, 1;" : "+r"(cnt[i]));                              // IMAD
            if (MODE == 2) asm volatile("mad.hi.u32 %0, %1, 2, %0;" : "+r"(cnt[i]) : "r"(aux[i]));               // IMAD.HI: cnt += aux >> 31
            if (MODE == 3) asm volatile("lop3.b32 %0, %0, 0x78, %1, 0xf8;" : "+r"(cnt[i]) : "r"(aux[i]));         // LOP3 (2 regs + imm)
            if (MODE == 4) asm volatile("shf.l.wrap.b32 %0, %0, %0, 3;" : "+r"(cnt[i]));                          // SHF
            if (MODE == 5) cnt[i] += aux[i] >> 31;                                                                  // compiler: LEA.HI
            if (MODE == 6) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tab_s + ((cnt[i] & 7u) << 3))); acc[i] += 0.0 * v; }   // (adds a DADD: see mode 9)
            if (MODE == 7) { double v; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(v) : "d"(acc[i])); cnt[i] ^= __double2hiint(v); }           // MUFU + LOP3
            if (MODE == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(cnt[i]) : "r"(aux[i]));                        // IADD3
            if (MODE == 9) acc[i] += 0.0 * a;                                                                        // the extra DADD/DFMA of mode 6 alone
            if (MODE == 10) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tab_s + ((unsigned)i << 3))); cnt[i] = v; }          // LDS only (fixed address)
            if (MODE == 11) asm volatile("mad.hi.u32 %0, %0, 128, 0;" : "+r"(cnt[i]));                               // IMAD.HI (1 reg)
            if (MODE == 12) asm volatile("shl.b32 %0, %0, 3;" : "+r"(cnt[i]));                                       // shift left (compiler picks IMAD.SHL / SHF)
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
    const char* names[] = {"DFMA alone", "+IMAD (mad.lo)", "+IMAD.HI (2 regs)", "+LOP3 (2 regs+imm)", "+SHF", "+LEA.HI (cnt += x>>31)", "+LDS.64 (var addr) +DADD",
                           "+MUFU.RCP64H +LOP3", "+IADD3 (2 regs)", "+DADD", "+LDS.32 fixed addr", "+IMAD.HI (1 reg)", "+shl (compiler)"};
    float ms;
    const double clk_per_ms = p.clockRate;
#define RUN(M) ms = run<M>(out, in, sms, iters); printf("%-28s %8.3f ms  %5.2f clk per group per sub-partition\n", names[M], ms, ms * clk_per_ms / (4.0 * 8 * iters));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    CK(cudaDeviceSynchronize());
    return 0;
}
