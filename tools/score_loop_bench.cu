// score_loop_bench.cu -- the phase-B hot loop of k_score / k_score_ws in isolation (16 warps per SM, tables in shared memory),
// with compile-time ablations, to find what bounds it:  nvcc -DVAR=<n> ...   prints clocks per (32 train points x 4 grid
// points) warp step per SM sub-partition (fp64 pipe alone: 44 instructions x 2 = 88).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared -DVAR=0 -o tools/score_loop_bench tools/score_loop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../conformal_bayesopt_b200/csrc/score.cuh"
#ifndef VAR
#define VAR 0
#endif
#ifndef NW
#define NW 16          // warps per SM (4 per sub-partition in the kernels)
#endif
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
using namespace cbo;

// (the loop k_score_ws used before score_rows_fx; kept here as VAR 0, the baseline of the comparison)
__constant__ double kK8[4];   // a2, a1, a0, c3 of the monic cubic
// ---- score_rows8: the same sweep, arranged for the REGISTER-FILE read ports -------------------------------------------
// Measured on B200 (tools/dfma_operands*.cu, profiles/r02_dfma_operands.log): a DFMA that reads three distinct register
// pairs issues every 3 clocks (66 % of the fp64 pipe), one that reads two (or one) every 2 clocks -- the pipe's own rate;
// immediates, constant-bank / uniform-register operands and operand-reuse-cache hits are free, and integer instructions
// pay for their register operands on the same ports.  score_rows spent 149 clocks per (32 train points x 4 grid points)
// against 88 of fp64-pipe time.  Here the cubic P(y) = exp(-y ln2/16) is evaluated in MONIC form,
//     P = c3 (((y + a2) y + a1) y + a0),   a_k = c_k / c3,
// so that every step has ONE constant operand (Horner starts with fma(y, c3, c2): two constants, i.e. two more register
// pairs per instruction once the constants sit in registers), and c3 is folded into the table: tab8[j] = c3 2^(j/16), so
// e = 2^(n/16) P(y) is still one fma on the exponent-scaled entry.  Same 11 fp64 instructions per element, ~20 register
// pairs read instead of 24.  (tools/score_loop_bench.cu times the loop alone with ablations.)
__device__ __forceinline__ void exp_tab8_init(double* tab, int j, double c3) {
    double v = c3 * exp2(double(j) / 16.0);
    tab[j] = __hiloint2double(__double2hiint(v) - (j << 16), __double2loint(v));
}

template <bool SOFT, bool CLAMP, bool UNIT, int GPW>
__device__ __forceinline__ void score_rows8(const ScoreArgs& p, const double* __restrict__ sq, const double* __restrict__ sdr,
                                            unsigned tab_s, int lane, int N32,
                                            const double (&bg)[GPW], double (&sr)[GPW], double (&sw)[GPW],
                                            unsigned (&cneg)[GPW]) {
    const double kMagic = p.k_magic;                          // 1.5 * 2^52
    const double kA2 = kK8[0], kA1 = kK8[1], kA0 = kK8[2];
    constexpr int kSatHi = 0x408d9000;                        // hiword(946.0)
    const double* tq = sq + 3 * lane;
    const double* td = sdr + lane;
    double qa = tq[0], qb = tq[1], qc = tq[2], dr = UNIT ? 1.0 : td[0];
    for (int i = lane; i < N32; i += 32) {
        tq += 96; td += 32;                                   // next row (reads past N32 stay inside shared memory)
        const double qa_n = tq[0], qb_n = tq[1], qc_n = tq[2], dr_n = UNIT ? 1.0 : td[0];
        double x[GPW];
#pragma unroll
        for (int k = 0; k < GPW; ++k) x[k] = fma(qa, bg[k], qb);
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            x[k] = fma(x[k], bg[k], qc);
            cneg[k] += unsigned(__double2hiint(x[k])) >> 31;  // counts score_c < score_i
        }
        if (SOFT) {
            double t[GPW], y[GPW], e[GPW];
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                double xc = x[k];
                if (CLAMP) {
                    int hi = __double2hiint(x[k]);
                    int ch = min(hi & 0x7fffffff, kSatHi) | (hi & 0x80000000);
                    xc = __hiloint2double(ch, __double2loint(x[k]));         // clamp(x, -946, 946)
                }
                t[k] = kMagic - xc;                           // low word = -round(x)
                y[k] = xc + (t[k] - kMagic);                  // x - round(x) in [-1/2, 1/2]
            }
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                double pl = y[k] + kA2;
                pl = fma(pl, y[k], kA1);
                pl = fma(pl, y[k], kA0);                      // exp(-y ln2/16) / c3
                double ts = exp_tab_scaled(tab_s, __double2loint(t[k]));     // c3 2^(j/16) 2^kk (tab_s: the tab8 table)
                e[k] = fma(ts, pl, 1.0);                      // 1 + e
            }
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                double d = e[k];
                double y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
                double err1 = fma(-d, y0, 2.0);               // Newton: rho = y0 (2 - d y0), folded into the running sum
                if (UNIT) {
                    sr[k] = fma(y0, err1, sr[k]);
                } else {
                    double rho = y0 * err1;
                    sr[k] += rho;
                    sw[k] = fma(rho, dr, sw[k]);
                }
            }
        } else if (!UNIT) {
#pragma unroll
            for (int k = 0; k < GPW; ++k) sw[k] += (__double2hiint(x[k]) >= 0) ? dr : 0.0;
        }
        qa = qa_n; qb = qb_n; qc = qc_n; dr = dr_n;
    }
}



template <int V>
__device__ __forceinline__ void rows(const ScoreArgs& p, const double* __restrict__ sq, unsigned tab_s, unsigned tab8_s, int lane, int N32,
                                     const double (&bg)[4], double (&sr)[4], unsigned (&cneg)[4]) {
    if (V == 20) {            // the round-1 loop
        double sw[4] = {0, 0, 0, 0};
        score_rows<true, false, true, 4>(p, sq, sq, tab_s, lane, N32, bg, sr, sw, cneg);
        return;
    }
    if (V == 0) {             // the loop the kernels use
        double sw[4] = {0, 0, 0, 0};
        score_rows8<true, false, true, 4>(p, sq, sq, tab8_s, lane, N32, bg, sr, sw, cneg);
        return;
    }
    if (V == 50) {            // score_rows_fx: exponent split folded into the quadratic, 8 fp64 instructions per element
        double sw[4] = {0, 0, 0, 0};
        score_rows_fx<false, true, 4>(p, sq, sq, tab_s, lane, N32, bg, sr, sw, cneg);
        return;
    }
    if (V == 52) {            // ... with a ratio estimator (10 fp64 instructions per element)
        double sw[4] = {0, 0, 0, 0};
        score_rows_fx<false, false, 4>(p, sq, sq, tab_s, lane, N32, bg, sr, sw, cneg);
        sr[0] += sw[0]; sr[1] += sw[1]; sr[2] += sw[2]; sr[3] += sw[3];
        return;
    }
    if (V == 53) {            // ... clamped
        double sw[4] = {0, 0, 0, 0};
        score_rows_fx<true, true, 4>(p, sq, sq, tab_s, lane, N32, bg, sr, sw, cneg);
        return;
    }
    if (V == 65) {   // score_rows_fx with ONE reciprocal per two grid points: 1/d0 = d1 r, 1/d1 = d0 r, r = 1/(d0 d1)
        const double kB2 = p.fx_b2, kB1 = p.fx_b1, kB0 = p.fx_b0;
        const double* tq = sq + 3 * lane;
        double qa = tq[0], qb = tq[1], qc = tq[2];
        const int nsteps = N32 >> 5;
        for (int it = 0; it < nsteps; ++it) {
            tq += 96;
            const double qa_n = tq[0], qb_n = tq[1], qc_n = tq[2];
            double t[4], e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = fma(qa, bg[k], qb);
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = fma(-t[k], bg[k], qc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned z = unsigned(__double2loint(t[k]));
                const int hi = __double2hiint(t[k]);
                cneg[k] += unsigned(hi) >> 19;
                const double w = __hiloint2double(int(0x3ff00000u + (z >> 12)), int(z << 20));
                double pl = w + kB2;
                pl = fma(pl, w, kB1);
                pl = fma(pl, w, kB0);
                e[k] = fma(exp_tab_scaled(tab_s, hi), pl, 1.0);
            }
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                const double m = e[k] * e[k + 1];
                double y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(m));
                const double r = y0 * fma(-m, y0, 2.0);
                sr[k] = fma(e[k + 1], r, sr[k]);
                sr[k + 1] = fma(e[k], r, sr[k + 1]);
            }
            qa = qa_n; qb = qb_n; qc = qc_n;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) cneg[k] -= unsigned(nsteps) * 0x826u;
        return;
    }
    if (V == 63 || V == 64) {   // score_rows_fx, software-pipelined: the quadratic of row i+1 is issued beside the sigmoid of row i
        const double kB2 = p.fx_b2, kB1 = p.fx_b1, kB0 = p.fx_b0;
        const double* tq = sq + 3 * lane;
        double t[4];
        {
            const double qa = tq[0], qb = tq[1], qc = tq[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = fma(-fma(qa, bg[k], qb), bg[k], qc);
        }
        const int nsteps = N32 >> 5;
        for (int it = 0; it < nsteps; ++it) {
            tq += 96;
            const double qa = tq[0], qb = tq[1], qc = tq[2];
            double tn[4], e[4];
            if (V == 63) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tn[k] = fma(qa, bg[k], qb);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned z = unsigned(__double2loint(t[k]));
                const int hi = __double2hiint(t[k]);
                cneg[k] += unsigned(hi) >> 19;
                const double w = __hiloint2double(int(0x3ff00000u + (z >> 12)), int(z << 20));
                double pl = w + kB2;
                if (V == 64) tn[k] = fma(qa, bg[k], qb);
                pl = fma(pl, w, kB1);
                pl = fma(pl, w, kB0);
                const double ts = exp_tab_scaled(tab_s, hi);
                if (V == 64) tn[k] = fma(-tn[k], bg[k], qc);
                e[k] = fma(ts, pl, 1.0);
            }
            if (V == 63) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tn[k] = fma(-tn[k], bg[k], qc);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double d = e[k], y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
                double err1 = fma(-d, y0, 2.0);
                sr[k] = fma(y0, err1, sr[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = tn[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) cneg[k] -= unsigned(nsteps) * 0x826u;
        return;
    }
    const double kMagic = p.k_magic, kA2 = kK8[0], kA1 = kK8[1], kA0 = kK8[2];
    const double* tq = sq + 3 * lane;
    double qa = tq[0], qb = tq[1], qc = tq[2];
    double bgu[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bgu[k] = (V == 30 || V == 31) ? __shfl_sync(0xffffffffu, bg[k], 0) : bg[k];
        if (V == 40 || V == 42) {     // warp-uniform by construction: REDUX writes a UNIFORM register
            const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(bg[k]));
            const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(bg[k]));
            bgu[k] = __hiloint2double((int)hi, (int)lo);
        }
    }
    const int nsteps = N32 >> 5;
    for (int it = 0; it < nsteps; ++it) {
        tq += 96;
        const double qa_n = tq[0], qb_n = tq[1], qc_n = tq[2];
        double x[4], t[4], y[4], e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = fma(qa, bgu[k], qb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x[k] = fma(x[k], bgu[k], qc);
            if (V != 1 && V != 5 && V != 31) cneg[k] += unsigned(__double2hiint(x[k])) >> 31;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            t[k] = kMagic - x[k];
            y[k] = x[k] + (t[k] - kMagic);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double pl = y[k] + kA2;
            pl = fma(pl, y[k], kA1);
            if (V != 41 && V != 42) pl = fma(pl, y[k], kA0);
            double ts;
            if (V == 41 || V == 42) {             // 128-entry table + quadratic (timing only: same constants)
                const int n = __double2loint(t[k]);
                unsigned addr = ((unsigned(n) << 3) & 0x3f8u) | tab_s;
                int lo, hi;
                asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
                ts = __hiloint2double(hi + (n << 13), lo);
            } else
            if (V == 2 || V == 5 || V == 31) ts = __hiloint2double(0xbee00000 + (__double2loint(t[k]) << 16), 0);   // no table load
            else ts = exp_tab_scaled(tab8_s, __double2loint(t[k]));
            e[k] = fma(ts, pl, 1.0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double d = e[k], y0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
            double err1 = fma(-d, y0, 2.0);
            sr[k] = fma(y0, err1, sr[k]);
        }
        qa = qa_n; qb = qb_n; qc = qc_n;
    }
}

__global__ void __launch_bounds__(NW * 32, 1) k(const ScoreArgs p, const double* coef, double* out, int passes) {
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(128) double tab8[16];
    __shared__ __align__(1024) double tab[128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) { exp_tab8_init(tab8, tid, kK8[3]); exp_tab_init(tab, tid); }
    if (tid >= 16 && tid < 128) tab[tid] = tab[tid & 15] ;
    if (VAR >= 50) {
        __syncthreads();
        if (tid < 16) exp_tabfx_init(tab, tid);
    }
    double* sq = sm + size_t(warp & 7) * (3 * 1024 + 96);
    for (int i = (warp < 8) ? lane : 3 * 1024 + 96; i < 3 * 1024 + 96; i += 32) {   // warps 0..7 fill table (warp & 7)
        double v = coef[i % (3 * 1024)];
        if ((VAR >= 50) && (i % 3) == 2) v = 1572864.0 - v;      // M - qc, M = 1.5 * 2^20
        sq[i] = v;
    }
    __syncthreads();
    double tot = 0.0;
    unsigned ctot = 0;
    for (int ps = 0; ps < passes; ++ps) {
        double bg[4], sr[4] = {0, 0, 0, 0};
        unsigned cneg[4] = {0, 0, 0, 0};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) bg[kk] = 0.01 * ((ps * 4 + kk + warp) % 64) - 0.3;
        rows<VAR>(p, sq, smem_u32(tab), smem_u32(tab8), lane, 1024, bg, sr, cneg);
        tot += sr[0] + sr[1] + sr[2] + sr[3];
        ctot += cneg[0] + cneg[1] + cneg[2] + cneg[3];
    }
    out[blockIdx.x * (NW * 32) + tid] = tot + ctot;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    ScoreArgs a; memset(&a, 0, sizeof(a));
    const double c0 = 0.9999999988534117, c1n = -0.043321698775061875, c2 = 0.0009384214835622084, c3n = -1.3551125681305332e-05;
    a.k_magic = 6755399441055744.0; a.k_c0 = c0; a.k_c1n = c1n; a.k_c2 = c2; a.k_c3n = c3n;
    a.fx_b2 = kFxB2; a.fx_b1 = kFxB1; a.fx_b0 = kFxB0;
    { const double k8[4] = {c2 / c3n, c1n / c3n, c0 / c3n, c3n}; CK(cudaMemcpyToSymbol(kK8, k8, sizeof(k8))); }
    static double h[3 * 1024];
    for (int i = 0; i < 1024; ++i) { h[3 * i] = -40.0 - 0.01 * i; h[3 * i + 1] = 3.0 + 0.001 * i; h[3 * i + 2] = 20.0 * ((i % 17) - 8); }
    double *coef, *out;
    CK(cudaMalloc(&coef, sizeof(h))); CK(cudaMemcpy(coef, h, sizeof(h), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, sizeof(double) * sms * NW * 32));
    const size_t smem = sizeof(double) * 8 * (3 * 1024 + 96);
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int passes = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, NW * 32, smem>>>(a, coef, out, passes);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<<<sms, NW * 32, smem>>>(a, coef, out, passes); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double hsum = 0; static double ho[512]; CK(cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost)); for (double v : ho) hsum += v;
    // per sub-partition: 4 warps x passes x 32 steps
    {   // host reference of thread 0's total: train points 0, 32, 64, ... x every pass' 4 grid points (warp 0)
        double ref = 0; long cnt = 0;
        for (int ps = 0; ps < passes; ++ps) for (int kk = 0; kk < 4; ++kk) {
            double bgv = 0.01 * ((ps * 4 + kk + 0) % 64) - 0.3;
            for (int i = 0; i < 1024; i += 32) {
                double x = fma(fma(h[3 * i], bgv, h[3 * i + 1]), bgv, h[3 * i + 2]);
                ref += 1.0 / (1.0 + exp(-x * 0.6931471805599453 / 16.0)); cnt += x < 0;
            }
        }
        printf("thread 0: device %.12e  host reference %.12e (rho sum %+.3e, count %ld)\n", ho[0], ref + cnt, ho[0] - ref - cnt, cnt);
    }
    printf("NW %d VAR %d: %.3f ms  %.1f clk per warp step per sub-partition  (checksum %.6e)\n", NW, VAR, best, best * pr.clockRate / (NW / 4.0 * passes * 32), hsum);
    return 0;
}
