// cbo_api.cu -- C ABI of libcbo (see include/cbo.h): workspace carving and kernel orchestration.
// No torch types, no allocation, everything on the caller's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "../../include/cbo.h"
#include "bounds.cuh"
#include <cmath>
#include "common.cuh"
#include "gemm_tri.cuh"
#include "rows.cuh"
#include "score.cuh"
#include "score_ws.cuh"
#include "score_bwd.cuh"
#include "score_q.cuh"
#include "mobo.cuh"
#include "nei.cuh"
#include "rows_bwd.cuh"
#include "ratio_mma.cuh"
#include "ratio_train.cuh"
#include "model_build.cuh"
#include "model_append.cuh"
#include "score_q_bwd.cuh"
#include "nei_bwd.cuh"

namespace {

thread_local char g_last_err[256] = "";

int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s", where, cudaGetErrorString(e));
    return CBO_ERR_CUDA;
}
#define CBO_CUDA(call)                                         \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);  \
    } while (0)

constexpr int64_t kDefaultChunk = 65536;   // candidates per internal pass: 1.07 GB of U/V rows at N=1024 (4.5 % faster than 16384: fewer launch tails)

// ---- optional per-kernel-class timing (CUDA events on the caller's stream; bench.py's roofline) ----
enum { T_PREP = 0, T_ROWS = 1, T_GEMM1 = 2, T_GEMM2 = 3, T_SCORE = 4, T_SCORE_BWD = 5, T_GEMM_BWD = 6, T_ROWS_BWD = 7, T_NCLASS = 8 };
struct Timing {
    std::atomic<bool> enabled{false};
    std::mutex mu;                             // calls from several host threads may record into the same lists
    std::vector<cudaEvent_t> ev[T_NCLASS];     // start/stop pairs
    std::vector<cudaEvent_t> pool;
    int64_t launches[T_NCLASS] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t get() {
        cudaEvent_t e;
        if (!pool.empty()) { e = pool.back(); pool.pop_back(); return e; }
        cudaEventCreate(&e);
        return e;
    }
} g_timing;
struct TimeScope {
    int cls; cudaStream_t s; bool on;
    TimeScope(int c, cudaStream_t st, int nlaunch) : cls(c), s(st), on(g_timing.enabled) {
        if (on) {
            std::lock_guard<std::mutex> l(g_timing.mu);
            cudaEvent_t e = g_timing.get(); cudaEventRecord(e, s); g_timing.ev[cls].push_back(e); g_timing.launches[cls] += nlaunch;
        }
    }
    ~TimeScope() {
        if (on) {
            std::lock_guard<std::mutex> l(g_timing.mu);
            cudaEvent_t e = g_timing.get(); cudaEventRecord(e, s); g_timing.ev[cls].push_back(e);
        }
    }
};
constexpr size_t kAlign = 256;

size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
    template <typename T>
    T* take(size_t n) {
        T* r = reinterpret_cast<T*>(base ? base + off : nullptr);
        off += align_up(n * sizeof(T));
        return r;
    }
};

struct Workspace {
    double *zp, *tmp, *T, *a, *h0, *dr_train, *sum_dr, *XsT, *eps_absmax;
    uint32_t* zb;
    double *UW, *V, *mu, *ut, *drc, *rn;
    double *cscale, *nacc, *acq_scratch, *stats;      // refinement only
    double* nei_values;                               // cNEI only: C x S x G
    double *val_bar, *cv_bar;                         // cNEI gradient only: C x S x G value adjoints, C x nb adjoints of c
    double* splitk;                                   // split-K partial tiles (calls with <= 32 candidate points)
    double* splitk2;                                  // ... of the probe solves, which run beside the candidates' on the latency path
    double* U2;                                       // latency path: second copy of the kernel rows for u^T T (runs beside pass 2)
    double *X3, *adj;                                 // gradient only: C x Np solve scratch, C x ADJ_STRIDE adjoints
    size_t bytes;
};

int64_t pick_chunk(int64_t b, const cbo_params* p, int64_t chunk) {
    if (p->max_grid_refinements > 0) return b;     // batch-global statistics need every candidate resident
    if (chunk <= 0) chunk = kDefaultChunk;
    return chunk < b ? chunk : b;
}

Workspace carve(void* ws, int64_t N, int64_t d, int64_t b, const cbo_params* p, int64_t chunk, bool grad = false) {
    const int64_t Np = cbo_padded_n(N), Sp = p->num_probes, C = pick_chunk(b, p, chunk) * p->q;   // rows per chunk
    const int64_t tiles_n = Np / cbo::GEMM_BN;
    Carver c(ws);
    Workspace w;
    w.zp = c.take<double>(Np * Sp);
    w.tmp = c.take<double>(Np * Sp);
    w.T = c.take<double>(Np * Sp);
    w.a = c.take<double>(Np);
    w.h0 = c.take<double>(Np);
    w.dr_train = c.take<double>(Np);
    w.sum_dr = c.take<double>(8);
    w.zb = c.take<uint32_t>(Np);
    w.XsT = c.take<double>(Np * d);
    w.eps_absmax = c.take<double>(p->eps_per_candidate ? b : 8);
    w.UW = c.take<double>(C * Np);
    w.V = c.take<double>(C * Np);
    w.mu = c.take<double>(C);
    w.ut = c.take<double>(C * Sp);
    w.drc = c.take<double>(C);
    w.rn = c.take<double>(C * tiles_n);
    const int64_t R = p->max_grid_refinements;
    w.cscale = c.take<double>(R > 0 ? (R + 1) * b * p->q : 1);
    w.nacc = c.take<double>(R > 0 ? b * p->q : 1);
    w.acq_scratch = c.take<double>(R > 0 ? b : 1);
    w.stats = c.take<double>(8);
    w.nei_values = c.take<double>((p->kind == CBO_ACQ_NEI || p->kind == CBO_ACQ_JOINT) ? C * p->num_mc * p->grid_res : 1);
    w.splitk = c.take<double>(cbo::gemm_splitk_scratch_doubles((int)Np));
    w.splitk2 = c.take<double>(cbo::gemm_splitk_scratch_doubles((int)Np));
    w.U2 = c.take<double>(int64_t(cbo::GEMM_SPLITK_MAX_M) * Np);
    const bool nei_kind = p->kind == CBO_ACQ_NEI || p->kind == CBO_ACQ_JOINT;
    w.val_bar = c.take<double>(grad && nei_kind ? C * p->num_mc * p->grid_res : 1);
    w.cv_bar = c.take<double>(grad && nei_kind ? C * cbo::NEI_MAX_NB : 1);
    w.X3 = c.take<double>(grad ? C * Np : 1);
    w.adj = c.take<double>(grad ? C * (cbo::ADJQ_STRIDE > cbo::ADJ_STRIDE ? cbo::ADJQ_STRIDE : cbo::ADJ_STRIDE) : 1);
    w.bytes = c.off;
    return w;
}

// min / max / sum((n/G - 1/2)^2) of the accepted counts, in float like the reference (helpers.py:150-155); the host divides
// by the batch size (after the optional cross-rank reduction)
__global__ void k_accept_stats(const double* __restrict__ nacc, int b, int G, double* __restrict__ stats) {
    __shared__ float smin[256], smax[256], ssum[256];
    float mn = 3.4e38f, mx = -3.4e38f, sm = 0.f;
    for (int i = threadIdx.x; i < b; i += blockDim.x) {
        float n = (float)nacc[i];
        mn = fminf(mn, n); mx = fmaxf(mx, n);
        float d = n / (float)G - 0.5f;
        sm += d * d;
    }
    smin[threadIdx.x] = mn; smax[threadIdx.x] = mx; ssum[threadIdx.x] = sm;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + o]);
            smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
            ssum[threadIdx.x] += ssum[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[0] = smin[0]; stats[1] = smax[0]; stats[2] = ssum[0]; }
}

// covar_scale = 1 + sign(diff) |diff|^(2/q) in float (helpers.py:153-167), accumulated into the grid covariance scale
__global__ void k_rescale(const double* __restrict__ nacc, const double* __restrict__ cur, double* __restrict__ next,
                          int b, int G, int q) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    float d = (float)nacc[i] / (float)G - 0.5f;
    float ad = fabsf(d);
    float pw = (q == 1) ? ad * ad : (q == 2 ? ad : (float)pow((double)ad, (double)(2.f / (float)q)));
    float sg = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
    next[i] = (double)(1.f + sg * pw) * cur[i];
}

int validate(const cbo_gp_f64* gp, const cbo_params* p, int64_t b) {
    if (!gp || !p || b <= 0) return CBO_ERR_INVALID_ARG;
    if (!gp->X || !gp->alpha || !gp->linv || !gp->linv_t || !gp->inv_lengthscale) return CBO_ERR_INVALID_ARG;
    if (gp->N <= 0 || gp->d <= 0 || gp->Np != cbo_padded_n(gp->N)) return CBO_ERR_INVALID_ARG;
    if (p->grid_res < 2) return CBO_ERR_INVALID_ARG;                       // helpers.py:113
    if (p->kind != CBO_ACQ_UCB && p->kind != CBO_ACQ_EI && p->kind != CBO_ACQ_NEI && p->kind != CBO_ACQ_JOINT) return CBO_ERR_INVALID_ARG;
    if (p->kind == CBO_ACQ_NEI || p->kind == CBO_ACQ_JOINT) {
        if (!p->nei || !p->nei->Xb || !p->nei->KbX || !p->nei->Sbb || !p->nei->mub || p->nei->nb < 1) return CBO_ERR_INVALID_ARG;
        if (p->nei->nb > cbo::NEI_MAX_NB || p->nei->nb + p->q > cbo::NEI_MAX_M) return CBO_ERR_UNSUPPORTED;
    }
    if (p->q < 1 || p->q > cbo::MAX_Q) return CBO_ERR_UNSUPPORTED;
    if (p->q > 1 && cbo::score_q_smem_bytes((int)gp->Np, (int)p->grid_res, (int)p->q) > 227 * 1024) return CBO_ERR_UNSUPPORTED;
    if (gp->d > cbo::MAX_D) return CBO_ERR_UNSUPPORTED;
    if (p->num_probes < 1 || p->num_probes > cbo::MAX_SP) return CBO_ERR_UNSUPPORTED;
    if (p->num_mc < 1 || p->num_mc > cbo::MAX_S) return CBO_ERR_UNSUPPORTED;
    if (p->grid_res > 4096) return CBO_ERR_UNSUPPORTED;
    if (cbo::score_pairs_for((int)gp->Np, (int)p->grid_res) < 1) return CBO_ERR_UNSUPPORTED;
    return CBO_OK;
}

std::atomic<int> g_num_sms[cbo::kMaxDevices];      // per device (zero-initialised)
int num_sms() {
    const int dev = cbo::current_device_slot();
    if (dev < 0) return 148;
    int n = g_num_sms[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_num_sms[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// CBO_SCORE_WS=0 / 1 forces the pair kernel / the warp-specialised kernel (A/B timing, tests of small batches); unset: by size
int score_ws_override() {
    const char* e = getenv("CBO_SCORE_WS");
    return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1;
}

template <bool SOFT>
cudaError_t launch_score(const cbo::ScoreArgs& a_in, cudaStream_t st, bool many) {
    cbo::ScoreArgs a = a_in;
    // a call with many candidates: warp-specialised kernel (4 candidates in flight per SM, helper warps for phases A/C/D)
    const int ov = score_ws_override();
    if ((ov == 1 || (ov < 0 && many && SOFT)) && cbo::score_ws_fits(a.Np, a.G)) {   // hard path: the sweep is too short for one helper warp per candidate to keep up
        const size_t smem = cbo::score_ws_smem_bytes(a.Np, a.G);
        static cbo::SmemOptIn optin_ws;
        if (cudaError_t e = optin_ws.ensure(cbo::k_score_ws<SOFT>, smem); e != cudaSuccess) return e;
        int grid = (a.n_cand + cbo::SWS_GROUPS - 1) / cbo::SWS_GROUPS;
        if (grid > num_sms()) grid = num_sms();
        CBO_LAUNCHED(), cbo::k_score_ws<SOFT><<<grid, cbo::SWS_THREADS, smem, st>>>(a);
        return cudaGetLastError();
    }
    // warps per candidate: pairs when there are enough candidates to fill the SMs, wider groups for small batches
    int wl = 1;
    while (wl < 4 && a.n_cand * (1 << wl) < 16 * num_sms()) ++wl;
    while (wl > 1 && (1 << wl) > 2 * cbo::SCORE_FWD_PAIRS) --wl;
    // large N: shared memory admits only a few candidates per SM -- widen the groups until the CTA has its 16 warps again
    while (wl < 4 && (2 << wl) <= 2 * cbo::SCORE_FWD_PAIRS &&
           cbo::score_groups_for(a.Np, a.G, 1 << wl) * (1 << wl) < 2 * cbo::SCORE_FWD_PAIRS &&
           cbo::score_groups_for(a.Np, a.G, 2 << wl) * (2 << wl) > cbo::score_groups_for(a.Np, a.G, 1 << wl) * (1 << wl))
        ++wl;
    const int ngroups = cbo::score_groups_for(a.Np, a.G, 1 << wl);
    if (ngroups < 1) return cudaErrorInvalidConfiguration;
    a.wpg_log2 = wl;
    size_t smem = cbo::score_smem_bytes(a.Np, a.G, ngroups);
    static cbo::SmemOptIn optin;
    if (cudaError_t e = optin.ensure(cbo::k_score<SOFT>, smem); e != cudaSuccess) return e;
    int grid = (a.n_cand + ngroups - 1) / ngroups;     // persistent: one CTA per SM, one warp group per candidate in flight
    if (grid > num_sms()) grid = num_sms();
    CBO_LAUNCHED(), cbo::k_score<SOFT><<<grid, ngroups * (32 << wl), smem, st>>>(a);
    return cudaGetLastError();
}

template <bool SOFT>
cudaError_t launch_score_bwd(const cbo::ScoreBwdArgs& a_in, cudaStream_t st) {
    cbo::ScoreBwdArgs a = a_in;
    // 4 warps per candidate when there are enough candidates to fill the SMs (the per-group tables are larger than the
    // forward kernel's: 4 groups x 4 warps), wider groups for small batches
    int wl = 2;
    while (wl < 4 && a.f.n_cand * (1 << wl) < 16 * num_sms()) ++wl;
    while (wl < 4 && a.f.G > 4 * (32 << wl)) ++wl;
    if (a.f.G > 4 * (32 << wl)) return cudaErrorInvalidConfiguration;
    while (wl < 4 && cbo::score_bwd_groups_for(a.f.Np, a.f.G, 1 << wl) * (1 << wl) < 16 &&
           cbo::score_bwd_groups_for(a.f.Np, a.f.G, 2 << wl) * (2 << wl) > cbo::score_bwd_groups_for(a.f.Np, a.f.G, 1 << wl) * (1 << wl))
        ++wl;
    const int ngroups = cbo::score_bwd_groups_for(a.f.Np, a.f.G, 1 << wl);
    if (ngroups < 1) return cudaErrorInvalidConfiguration;
    a.f.wpg_log2 = wl;
    size_t smem = cbo::score_bwd_smem_bytes(a.f.Np, a.f.G, ngroups);
    static cbo::SmemOptIn optin;
    if (cudaError_t e = optin.ensure(cbo::k_score_bwd<SOFT>, smem); e != cudaSuccess) return e;
    int grid = (a.f.n_cand + ngroups - 1) / ngroups;
    if (grid > num_sms()) grid = num_sms();
    CBO_LAUNCHED(), cbo::k_score_bwd<SOFT><<<grid, ngroups * (32 << wl), smem, st>>>(a);
    return cudaGetLastError();
}

template <int Q, bool SOFT>
cudaError_t launch_score_q_bwd_t(const cbo::ScoreQBwdArgs& a, cudaStream_t st) {
    size_t smem = cbo::score_q_bwd_smem_bytes(a.f.base.Np, a.f.base.G, Q);
    // a handful of joint batches (the SGLD loop): one thread-block cluster of 8 CTAs per joint batch (soft path)
    if (SOFT && a.f.base.n_cand * cbo::SQB_CLUSTER <= num_sms()) {
        static cbo::SmemOptIn optin_cl;      // aux: clusters of 8 that are resident at once (GPC granularity), per smem size
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(a.f.base.n_cand * cbo::SQB_CLUSTER);
        cfg.blockDim = dim3(cbo::SQ_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cbo::SQB_CLUSTER; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        bool fresh = false;
        if (cudaError_t e = optin_cl.ensure(cbo::k_score_q_bwd<Q, SOFT, true>, smem, &fresh); e != cudaSuccess) return e;
        const int dev = cbo::current_device_slot() < 0 ? 0 : cbo::current_device_slot();
        if (fresh) {
            int mc = 0;
            if (cudaOccupancyMaxActiveClusters(&mc, cbo::k_score_q_bwd<Q, SOFT, true>, &cfg) != cudaSuccess) {
                cudaGetLastError();
                mc = 0;
            }
            optin_cl.aux[dev].store(mc, std::memory_order_release);
        }
        if (a.f.base.n_cand <= optin_cl.aux[dev].load(std::memory_order_acquire))   // one wave of clusters; otherwise one CTA per joint batch is as fast
            return CBO_LAUNCHED(), cudaLaunchKernelEx(&cfg, cbo::k_score_q_bwd<Q, SOFT, true>, a);
    }
    static cbo::SmemOptIn optin;
    if (cudaError_t e = optin.ensure(cbo::k_score_q_bwd<Q, SOFT, false>, smem); e != cudaSuccess) return e;
    int grid = a.f.base.n_cand < 2 * num_sms() ? a.f.base.n_cand : 2 * num_sms();
    CBO_LAUNCHED(), cbo::k_score_q_bwd<Q, SOFT, false><<<grid, cbo::SQ_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_score_q_bwd(const cbo::ScoreQBwdArgs& a, int q, bool soft, cudaStream_t st) {
    switch (q) {
        case 1: return soft ? launch_score_q_bwd_t<1, true>(a, st) : launch_score_q_bwd_t<1, false>(a, st);
        case 2: return soft ? launch_score_q_bwd_t<2, true>(a, st) : launch_score_q_bwd_t<2, false>(a, st);
        case 3: return soft ? launch_score_q_bwd_t<3, true>(a, st) : launch_score_q_bwd_t<3, false>(a, st);
        case 4: return soft ? launch_score_q_bwd_t<4, true>(a, st) : launch_score_q_bwd_t<4, false>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int Q, bool SOFT>
cudaError_t launch_score_q_t(const cbo::ScoreQArgs& a, cudaStream_t st) {
    size_t smem = cbo::score_q_smem_bytes(a.base.Np, a.base.G, Q);
    static cbo::SmemOptIn optin;
    if (cudaError_t e = optin.ensure(cbo::k_score_q<Q, SOFT>, smem); e != cudaSuccess) return e;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    int grid = a.base.n_cand < per_sm * num_sms() ? a.base.n_cand : per_sm * num_sms();
    // a handful of joint batches: 512 threads each (latency); many: 256 threads, more CTAs per SM
    const int threads = (Q <= 2 && a.base.n_cand < 2 * num_sms()) ? cbo::SQ_THREADS_SMALL : cbo::SQ_THREADS;
    CBO_LAUNCHED(), cbo::k_score_q<Q, SOFT><<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_score_q(const cbo::ScoreQArgs& a, int q, bool soft, cudaStream_t st) {
    switch (q) {
        case 2: return soft ? launch_score_q_t<2, true>(a, st) : launch_score_q_t<2, false>(a, st);
        case 3: return soft ? launch_score_q_t<3, true>(a, st) : launch_score_q_t<3, false>(a, st);
        case 4: return soft ? launch_score_q_t<4, true>(a, st) : launch_score_q_t<4, false>(a, st);
        default: return cudaErrorInvalidValue;
    }
}

template <int Q, bool BWD>
cudaError_t launch_ehvi_q(const cbo::EhviQArgs& a, size_t smem, cudaStream_t st) {
    static cbo::SmemOptIn optin;
    if (cudaError_t e = optin.ensure(cbo::k_ehvi_q<Q, BWD>, smem); e != cudaSuccess) return e;
    CBO_LAUNCHED(), cbo::k_ehvi_q<Q, BWD><<<(a.b + cbo::EHVIQ_WARPS - 1) / cbo::EHVIQ_WARPS, cbo::EHVIQ_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

int cehvi_integrate_q_impl(int64_t m, int64_t b, int64_t G, int64_t S, int64_t q, const double* const* mask,
                           const double* const* logp, const double* const* grid, const double* const* mu,
                           const double* const* ood, const double* const* cov, const double* noise_host,
                           const double* const* fx, const double* eps_mc, const double* cell_lower,
                           const double* cell_upper, int64_t n_cells, int32_t cells_per_sample, double alpha,
                           int32_t optimistic, double* acq, double* values_out, bool bwd, double* const* mask_bar,
                           double* const* ood_bar, double* const* f_bar, void* stream) {
    if (!mask || !logp || !grid || !mu || !ood || !cell_lower || !cell_upper || !acq) return CBO_ERR_INVALID_ARG;
    if (m < 1 || b <= 0 || G < 2 || S < 1 || n_cells < 1 || q < 1) return CBO_ERR_INVALID_ARG;
    if (!fx && (!cov || !noise_host || !eps_mc)) return CBO_ERR_INVALID_ARG;
    if (bwd && (!values_out || !mask_bar || !ood_bar || !f_bar)) return CBO_ERR_INVALID_ARG;
    if (m > cbo::MAX_M || q > cbo::MAX_Q) return CBO_ERR_UNSUPPORTED;
    const int64_t ncell_tot = n_cells * (cells_per_sample ? S : 1);
    const size_t smem = cbo::ehvi_q_smem_bytes((int)m, (int)q, (int)ncell_tot);
    if (smem > 200 * 1024) return CBO_ERR_UNSUPPORTED;
    cbo::EhviQArgs a;
    memset(&a, 0, sizeof(a));
    a.m = (int)m; a.b = (int)b; a.G = (int)G; a.S = (int)S; a.q = (int)q; a.n_cells = (int)n_cells; a.optimistic = optimistic;
    a.cells_per_sample = cells_per_sample; a.alpha_lvl = alpha;
    for (int k = 0; k < m; ++k) {
        if (!mask[k] || !logp[k] || !grid[k] || !mu[k] || !ood[k]) return CBO_ERR_INVALID_ARG;
        a.mask[k] = mask[k]; a.logp[k] = logp[k]; a.grid[k] = grid[k]; a.mu[k] = mu[k]; a.ood[k] = ood[k];
        a.fx[k] = fx ? fx[k] : nullptr;
        if (!a.fx[k]) {
            if (!cov || !cov[k] || !noise_host || !eps_mc) return CBO_ERR_INVALID_ARG;
            a.cov[k] = cov[k]; a.noise[k] = noise_host[k];
        }
        if (bwd) {
            if (!mask_bar[k] || !ood_bar[k]) return CBO_ERR_INVALID_ARG;
            a.mask_bar[k] = mask_bar[k]; a.ood_bar[k] = ood_bar[k]; a.f_bar[k] = f_bar[k];
        }
    }
    a.eps_mc = eps_mc; a.cell_lower = cell_lower; a.cell_upper = cell_upper; a.acq = acq; a.values = values_out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    switch (q) {
        case 1: e = bwd ? launch_ehvi_q<1, true>(a, smem, st) : launch_ehvi_q<1, false>(a, smem, st); break;
        case 2: e = bwd ? launch_ehvi_q<2, true>(a, smem, st) : launch_ehvi_q<2, false>(a, smem, st); break;
        case 3: e = bwd ? launch_ehvi_q<3, true>(a, smem, st) : launch_ehvi_q<3, false>(a, smem, st); break;
        default: e = bwd ? launch_ehvi_q<4, true>(a, smem, st) : launch_ehvi_q<4, false>(a, smem, st); break;
    }
    CBO_CUDA(e);
    return CBO_OK;
}

}  // namespace

extern "C" {

int cbo_version(void) { return CBO_VERSION; }

int64_t cbo_launch_count(void) { return (int64_t)cbo::launch_counter().load(std::memory_order_relaxed); }

#ifdef CBO_SQB_CLOCKS   // experiment builds only (tools/sqb_phase_clocks.py): phase clocks of CTA 0 of the last k_score_q_bwd
int cbo_debug_sqb_clocks(long long* out32) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out32, cbo::g_sqb_clk, sizeof(long long) * 32);
}
#endif

const char* cbo_status_str(int s) {
    switch (s) {
        case CBO_OK: return "ok";
        case CBO_ERR_INVALID_ARG: return "invalid argument";
        case CBO_ERR_UNSUPPORTED: return "unsupported shape";
        case CBO_ERR_WORKSPACE: return "workspace too small or misaligned";
        case CBO_ERR_CUDA: return "CUDA error";
        case CBO_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

const char* cbo_last_cuda_error(void) { return g_last_err; }

int64_t cbo_padded_n(int64_t N) { return (N + 127) / 128 * 128; }

size_t cbo_workspace_bytes(int64_t N, int64_t d, int64_t b, const cbo_params* params, int64_t chunk) {
    if (!params || N <= 0 || b <= 0) return 0;
    return carve(nullptr, N, d, b, params, chunk).bytes;
}

int cbo_gemm_tn_f64(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                    int64_t Nn, int64_t K, int32_t tri, double* rownorm_partial, void* stream) {
    if (!A || !B || !C || M <= 0 || Nn <= 0 || K <= 0) return CBO_ERR_INVALID_ARG;
    if (Nn % cbo::GEMM_BN || K % cbo::GEMM_BN || tri < 0 || tri > 2) return CBO_ERR_UNSUPPORTED;
    if ((lda | ldb | ldc) & 1) return CBO_ERR_UNSUPPORTED;      // 16-byte cp.async / double2 stores
    CBO_CUDA(cbo::launch_gemm_tn(A, lda, B, ldb, C, ldc, (int)M, (int)Nn, (int)K, tri, rownorm_partial,
                                 static_cast<cudaStream_t>(stream), num_sms()));
    return CBO_OK;
}

int cbo_kernel_rows_f64(const cbo_gp_f64* gp, const double* Xq, int64_t n, double* U, double* mu, double* scratch,
                        void* stream) {
    if (!gp || !Xq || !U || !scratch || n <= 0) return CBO_ERR_INVALID_ARG;
    if (gp->d > cbo::MAX_D) return CBO_ERR_UNSUPPORTED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CBO_LAUNCHED(), cbo::k_scale_transpose<<<(int)((gp->Np * gp->d + 255) / 256), 256, 0, s>>>(gp->X, (int)gp->N, (int)gp->Np, (int)gp->d,
                                                                             gp->inv_lengthscale, scratch);
    CBO_CUDA(cbo::launch_kernel_rows(Xq, (int)n, scratch, (int)gp->N, (int)gp->Np, (int)gp->d, gp->inv_lengthscale,
                                     gp->outputscale, gp->mean, gp->alpha, nullptr, 0, U, mu, nullptr, num_sms(), s));
    CBO_CUDA(cudaGetLastError());
    return CBO_OK;
}

int cbo_ratio_forward_f64(const cbo_ratio_mlp_f64* r, const double* x, int64_t n, int64_t d, double* out, void* stream) {
    if (!r || !x || !out || n <= 0 || d <= 0) return CBO_ERR_INVALID_ARG;
    if (r->hidden < 1 || r->hidden > cbo::MAX_H) return CBO_ERR_UNSUPPORTED;
    if (d > cbo::MAX_D) return CBO_ERR_UNSUPPORTED;
    CBO_CUDA(cbo::launch_ratio_mlp(x, (int)n, (int)d, (int)r->hidden, r->W1, r->b1, r->W2, r->b2, r->W3, r->b3, 0.0, out,
                                   num_sms(), static_cast<cudaStream_t>(stream)));
    return CBO_OK;
}

size_t cbo_ratio_train_workspace_bytes(int64_t n_rows, int64_t d) {
    if (n_rows <= 0 || d <= 0 || d > cbo::RT_MAX_D) return 0;
    const int64_t tiles = (n_rows + cbo::RT_ROWS - 1) / cbo::RT_ROWS;
    return size_t(tiles) * size_t(cbo::ratio_num_params((int)d) + 1) * sizeof(double);
}

int cbo_ratio_train_step_f64(int64_t d, int64_t hidden, double* theta, double* adam_m, double* adam_v, double* target,
                             const double* x_neg, int64_t n_neg, const double* x_pos, const int64_t* pos_idx,
                             int64_t n_pos, int64_t step, double lr, double beta1, double beta2, double eps,
                             double weight_decay, double ema_weight, double* loss, void* workspace,
                             size_t workspace_bytes, void* stream) {
    if (!theta || !adam_m || !adam_v || !target || !workspace || d <= 0 || step < 1 || n_neg < 0 || n_pos < 0 ||
        n_neg + n_pos <= 0 || (n_neg > 0 && !x_neg) || (n_pos > 0 && !x_pos))
        return CBO_ERR_INVALID_ARG;
    if (hidden != cbo::RT_H || d > cbo::RT_MAX_D || n_neg + n_pos > (int64_t(1) << 30)) return CBO_ERR_UNSUPPORTED;
    const int64_t n_rows = n_neg + n_pos;
    if (workspace_bytes < cbo_ratio_train_workspace_bytes(n_rows, d) || (reinterpret_cast<uintptr_t>(workspace) & 7))
        return CBO_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles = int((n_rows + cbo::RT_ROWS - 1) / cbo::RT_ROWS);
    const size_t smem = cbo::ratio_train_smem_bytes((int)d);
    static cbo::SmemOptIn optin;
    CBO_CUDA(optin.ensure(cbo::k_ratio_train_grad, smem));
    double* partial = static_cast<double*>(workspace);
    CBO_LAUNCHED(), cbo::k_ratio_train_grad<<<tiles, cbo::RT_THREADS, smem, st>>>(theta, (int)d, x_neg, (int)n_neg, x_pos, pos_idx,
                                                                  (int)n_pos, partial);
    CBO_CUDA(cudaGetLastError());
    const int64_t P = cbo::ratio_num_params((int)d);
    const double bc1 = 1.0 - std::pow(beta1, double(step)), bc2 = 1.0 - std::pow(beta2, double(step));
    CBO_LAUNCHED(), cbo::k_ratio_adam_ema<<<int((P + 1 + 255) / 256), 256, 0, st>>>(theta, adam_m, adam_v, target, partial, tiles, P,
                                                                   double(n_rows), lr / bc1, 1.0 / std::sqrt(bc2), beta1,
                                                                   beta2, eps, weight_decay, ema_weight, loss);
    CBO_CUDA(cudaGetLastError());
    return CBO_OK;
}

int cbo_sample_subset_i64(int64_t n, int64_t k, uint64_t seed, int64_t* out_idx, void* stream) {
    if (n <= 0 || k < 0 || k > n || (k > 0 && !out_idx)) return CBO_ERR_INVALID_ARG;
    if (n > (int64_t(1) << 31)) return CBO_ERR_UNSUPPORTED;
    if (k == 0) return CBO_OK;
    int half_bits = 1;
    while ((int64_t(1) << (2 * half_bits)) < n) ++half_bits;
    CBO_LAUNCHED(), cbo::k_sample_subset<<<int((k + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, k, seed, half_bits, out_idx);
    CBO_CUDA(cudaGetLastError());
    return CBO_OK;
}

int cbo_sgld_update_f64(double* X, double* momentum_buf, const double* dacq_dX, const double* noise, const double* lb,
                        const double* ub, int64_t n_rows, int64_t d, double lr, double momentum, double temperature,
                        double penalty_scale, double penalty_width, void* stream) {
    if (!X || !dacq_dX || !noise || !lb || !ub || n_rows <= 0 || d <= 0 || lr < 0.0 || momentum < 0.0 ||
        temperature < 0.0 || penalty_width <= 0.0 || (momentum > 0.0 && !momentum_buf))
        return CBO_ERR_INVALID_ARG;
    const int64_t n = n_rows * d;
    const double ns = momentum > 0.0 ? std::sqrt(2.0 * lr * momentum * temperature) : std::sqrt(2.0 * lr * temperature);
    CBO_LAUNCHED(), cbo::k_sgld_update<<<int((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        X, momentum_buf, dacq_dX, noise, lb, ub, n, (int)d, lr, momentum, ns, penalty_scale, penalty_width);
    CBO_CUDA(cudaGetLastError());
    return CBO_OK;
}

size_t cbo_model_append_workspace_bytes(int64_t Np) {
    return Np > 0 ? cbo::model_append_ws_doubles(Np) * sizeof(double) : 0;
}

int cbo_model_append_f64(double* X, int64_t N, int64_t d, int64_t Np, const double* inv_lengthscale, double mean,
                         double outputscale, double noise, double* alpha, double* linv, double* linv_t,
                         const double* x_new, const double* y_new, int64_t q, int32_t* info, void* workspace,
                         size_t workspace_bytes, void* stream) {
    if (!X || !inv_lengthscale || !alpha || !linv || !linv_t || !x_new || !y_new || !workspace || N < 0 || d <= 0 || q <= 0)
        return CBO_ERR_INVALID_ARG;
    if (d > cbo::MAX_D || Np % 128 != 0 || N + q > Np) return CBO_ERR_UNSUPPORTED;
    if (workspace_bytes < cbo_model_append_workspace_bytes(Np) || (reinterpret_cast<uintptr_t>(workspace) & 7))
        return CBO_ERR_WORKSPACE;
    CBO_CUDA(cbo::model_append(X, N, d, Np, inv_lengthscale, mean, outputscale, noise, alpha, linv, linv_t, x_new, y_new, q,
                               info, static_cast<double*>(workspace), static_cast<cudaStream_t>(stream)));
    return CBO_OK;
}

int cbo_conf_mask_to_bounds_f64(const double* grid, const double* mask, int64_t b, int64_t G, int64_t t, double* lb,
                                double* ub, void* stream) {
    if (!grid || !mask || !lb || !ub || b <= 0 || G <= 0 || t <= 0) return CBO_ERR_INVALID_ARG;
    if (G > cbo::BOUNDS_MAX_G) return CBO_ERR_UNSUPPORTED;
    int Gp = 2;
    while (Gp < G) Gp <<= 1;
    CBO_LAUNCHED(), cbo::k_mask_to_bounds<<<(int)(b * t), cbo::BOUNDS_THREADS, 2 * Gp * sizeof(double),
                            static_cast<cudaStream_t>(stream)>>>(grid, mask, (int)b, (int)G, (int)t, Gp, lb, ub);
    CBO_CUDA(cudaGetLastError());
    return CBO_OK;
}

// Side streams of the latency path (a call with a handful of candidate points, e.g. one SGLD step of 8 restarts x q = 2).
// Such a call is a chain of ~20 kernels of 3-15 us; three sub-chains do not depend on each other until late:
//   caller's stream   scaled inputs -> kernel rows (U, mu) -> pass 1 (V, ||v||^2) -> [T] u^T T -> pass 2 (W) -> [tables, drc] scores ...
//   st2 (probes)      padded probes -> two split-K passes (T = A^-1 z) -> [dr_train] per-train-point tables
//   st  (ratio)       density-ratio MLP over the train points (dr_train), over the candidates (drc); in the reverse pass the
//                     estimator's reverse beside the two reverse solves
// Fork / join through events only: legal under stream capture (the side streams become branches of the caller's graph), and
// the caller still sees all work ordered on its own stream when the call returns.  Two streams and five timing-free events
// per device, created on first use; `call_mu` serialises the enqueue of calls that share them.
struct SideStream {
    std::mutex mu, call_mu;
    cudaStream_t st = nullptr, st2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr, ev_T = nullptr, ev_tables = nullptr, ev_rows = nullptr, ev_ut = nullptr;
    cudaError_t ensure() {
        std::lock_guard<std::mutex> l(mu);
        if (st) return cudaSuccess;
        cudaStream_t x, y;
        if (cudaError_t e = cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking); e != cudaSuccess) return e;
        if (cudaError_t e = cudaStreamCreateWithFlags(&y, cudaStreamNonBlocking); e != cudaSuccess) return e;
        for (cudaEvent_t* ev : {&fork, &join, &join2, &ev_T, &ev_tables, &ev_rows, &ev_ut})
            if (cudaError_t e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming); e != cudaSuccess) return e;
        st2 = y;
        st = x;
        return cudaSuccess;
    }
};
SideStream g_side[cbo::kMaxDevices];

// Per-chunk callbacks of the host-buffer entry point: `before` is called ahead of the first kernel that reads the chunk's
// candidates, `after` behind the last kernel that writes its acquisition values (both only enqueue stream operations).
struct ChunkHooks {
    void* ctx;
    int (*before)(void* ctx, int64_t c0, int64_t ncand);
    int (*after)(void* ctx, int64_t c0, int64_t ncand);
};

static int conformal_acq_impl(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* p, const double* Xq,
                              int64_t b, const double* eps_grid, const double* z, const double* eps_mc,
                              const double* u_rand, const cbo_outputs_f64* out, double* dX, void* workspace,
                              size_t workspace_bytes, int64_t chunk, void* stream, const ChunkHooks* hooks = nullptr) {
    int st = validate(gp, p, b);
    if (st != CBO_OK) return st;
    const bool grad = dX != nullptr;
    const bool nei_kind = p->kind == CBO_ACQ_NEI || p->kind == CBO_ACQ_JOINT;
    const cbo_ext_adjoints_f64* ext = grad ? p->ext_adjoints : nullptr;
    // reverse pass: q = 1 cUCB / cEI on the throughput kernels; joint batches, cNEI / joint samples and external adjoints on
    // k_score_q_bwd (one CTA or cluster per joint batch, q = 1 included)
    const bool qpath = p->q > 1 || (grad && (nei_kind || ext));
    if (grad) {
        if (ext && (!ext->mask_bar || !ext->ood_bar)) return CBO_ERR_INVALID_ARG;
        if (p->kind == CBO_ACQ_JOINT && !ext) return CBO_ERR_UNSUPPORTED;     // joint samples have no integration of their own to differentiate
        if (!qpath && (p->grid_res > 2048 || cbo::score_bwd_groups_for((int)gp->Np, (int)p->grid_res, 16) < 1)) return CBO_ERR_UNSUPPORTED;
        if (qpath && cbo::score_q_bwd_smem_bytes((int)gp->Np, (int)p->grid_res, (int)p->q) > 227 * 1024) return CBO_ERR_UNSUPPORTED;
    }
    if (!Xq || !eps_grid || !z || !eps_mc || !out || !out->acq || !workspace) return CBO_ERR_INVALID_ARG;
    if (p->randomized && !u_rand) return CBO_ERR_INVALID_ARG;
    if (ratio && (ratio->hidden < 1 || ratio->hidden > cbo::MAX_H)) return CBO_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return CBO_ERR_WORKSPACE;
    Workspace w = carve(workspace, gp->N, gp->d, b, p, chunk, grad);
    if (w.bytes > workspace_bytes) return CBO_ERR_WORKSPACE;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool soft_path = p->temp > 0.0;
    const int N = (int)gp->N, Np = (int)gp->Np, d = (int)gp->d, Sp = (int)p->num_probes, G = (int)p->grid_res;
    const int tiles_n = Np / cbo::GEMM_BN;
    const int64_t C = pick_chunk(b, p, chunk);

    // latency path (one chunk of <= 32 candidate points): three branches, see SideStream
    SideStream* side = nullptr;
    std::unique_lock<std::mutex> side_lock;
    const char* side_env = getenv("CBO_SIDE_STREAM");                 // "0": everything on the caller's stream (A/B timing, tests)
    const bool side_enabled = !(side_env && side_env[0] == '0');
    // (not with grid refinement: those calls synchronise with the host and run the caller's reduce hook once per pass --
    // `call_mu` must never be held across a callback -- and their latency is dominated by that anyway)
    if (side_enabled && b * p->q <= cbo::GEMM_SPLITK_MAX_M && C >= b && !hooks && p->max_grid_refinements == 0) {
        const int dev = cbo::current_device_slot();
        if (dev >= 0) {
            side = &g_side[dev];
            CBO_CUDA(side->ensure());
            side_lock = std::unique_lock<std::mutex>(side->call_mu);
            CBO_CUDA(cudaEventRecord(side->fork, s));
            CBO_CUDA(cudaStreamWaitEvent(side->st2, side->fork, 0));
            CBO_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));      // (u^T T runs there also without a ratio estimator)
        }
    }
    cudaStream_t s_ratio = side ? side->st : s;       // density-ratio MLP launches
    cudaStream_t s_probe = side ? side->st2 : s;      // probe solves and per-train-point tables

    // ---- per-call tables: T = A^-1 z[:N], dr_train, a, h0, zb, sum_dr -------------------------------
    {
    TimeScope ts_(T_PREP, s, 7);
    CBO_LAUNCHED(), cbo::k_scale_transpose<<<(Np * d + 255) / 256, 256, 0, s>>>(gp->X, N, Np, d, gp->inv_lengthscale, w.XsT);
    CBO_LAUNCHED(), cbo::k_pad_probes<<<(Np * Sp + 255) / 256, 256, 0, s_probe>>>(z, N, Np, Sp, w.zp);
    // T^T = Z^T A^-1 (Sp x Np): the probe vectors are Sp more "rows" for the two triangular passes (split-K: a handful of rows)
    double* splitk_probe = side ? w.splitk2 : w.splitk;      // the candidates' split-K passes may run at the same time
    CBO_CUDA(cbo::launch_gemm_tn(w.zp, Np, gp->linv, Np, w.tmp, Np, Sp, Np, Np, cbo::TRI_LOWER, nullptr, s_probe, num_sms(), splitk_probe));
    CBO_CUDA(cbo::launch_gemm_tn(w.tmp, Np, gp->linv_t, Np, w.T, Np, Sp, Np, Np, cbo::TRI_UPPER, nullptr, s_probe, num_sms(), splitk_probe));
    if (side) CBO_CUDA(cudaEventRecord(side->ev_T, side->st2));
    if (ratio) {
        CBO_CUDA(cbo::launch_ratio_mlp(gp->X, N, d, (int)ratio->hidden, ratio->W1, ratio->b1, ratio->W2, ratio->b2,
                                       ratio->W3, ratio->b3, 1e-6, w.dr_train, num_sms(), s_ratio));
        if (side) {                                            // k_train_tables below reads dr_train
            CBO_CUDA(cudaEventRecord(side->join, side->st));
            CBO_CUDA(cudaStreamWaitEvent(side->st2, side->join, 0));
        }
    } else
        CBO_LAUNCHED(), cbo::k_fill<<<(Np + 255) / 256, 256, 0, s_probe>>>(w.dr_train, Np, 1.0);
    if (p->q == 1) {
        const int rows = p->eps_per_candidate ? (int)b : 1;
        CBO_LAUNCHED(), cbo::k_eps_absmax<<<(rows + 7) / 8, 256, 0, s_probe>>>(eps_grid, rows, G, w.eps_absmax);
    }
    CBO_LAUNCHED(), cbo::k_train_tables<<<1, 256, 0, s_probe>>>(gp->alpha, z, w.T, N, Np, Sp, gp->noise, w.dr_train, w.a, w.h0, w.zb,
                                          w.sum_dr);
    if (side) CBO_CUDA(cudaEventRecord(side->ev_tables, side->st2));
    }
    CBO_CUDA(cudaGetLastError());

    const int R = (int)p->max_grid_refinements;
    const int q = (int)p->q;
    const bool many = b >= int64_t(2) * cbo::SWS_GROUPS * num_sms();   // decided per CALL (chunking stays bit-invariant): every group gets >= 2 candidates
    double* splitk = (b * q <= cbo::GEMM_SPLITK_MAX_M) ? w.splitk : nullptr;   // a handful of candidate points in the whole call: split-K GEMMs
    for (int64_t c0 = 0; c0 < b; c0 += C) {
        const int ncand = (int)((b - c0 < C) ? (b - c0) : C);
        const int n = ncand * q;                                  // kernel rows of this chunk: one per candidate POINT
        const double* Xc = Xq + c0 * q * d;
        if (hooks) { int hs_ = hooks->before(hooks->ctx, c0, ncand); if (hs_ != CBO_OK) return hs_; }
        // ---- kernel rows, mu, u^T T; candidate density ratios -------------------------------------
        {
            TimeScope ts_(T_ROWS, s, ratio ? 2 : 1);
            // latency path: the rows do not wait for T (the probe branch); u^T T follows between the two passes
            CBO_CUDA(cbo::launch_kernel_rows(Xc, n, w.XsT, N, Np, d, gp->inv_lengthscale, gp->outputscale, gp->mean,
                                             gp->alpha, side ? nullptr : w.T, Sp, w.UW, w.mu, side ? nullptr : w.ut, num_sms(), s,
                                             splitk != nullptr, side ? w.U2 : nullptr));
            if (side) CBO_CUDA(cudaEventRecord(side->ev_rows, s));
            if (ratio)
                CBO_CUDA(cbo::launch_ratio_mlp(Xc, n, d, (int)ratio->hidden, ratio->W1, ratio->b1, ratio->W2, ratio->b2,
                                               ratio->W3, ratio->b3, 1e-6, w.drc, num_sms(), s_ratio));
            if (side && ratio) CBO_CUDA(cudaEventRecord(side->join2, side->st));
            CBO_CUDA(cudaGetLastError());
        }
        // ---- v = L^-1 u (+ ||v||^2 partials), w = L^-T v ------------------------------------------
        {
            TimeScope ts_(T_GEMM1, s, 1);
            CBO_CUDA(cbo::launch_gemm_tn(w.UW, Np, gp->linv, Np, w.V, Np, n, Np, Np, cbo::TRI_LOWER, w.rn, s, num_sms(), splitk));
        }
        if (side) {   // u^T T from a second copy of the stored rows, on the third stream beside pass 2 (which overwrites the rows);
                      // same accumulation order as the fused kernel: same bits
            CBO_CUDA(cudaStreamWaitEvent(side->st, side->ev_rows, 0));
            CBO_CUDA(cudaStreamWaitEvent(side->st, side->ev_T, 0));
            CBO_LAUNCHED(), cbo::k_ut_small<<<n, cbo::ROWS_SMALL_THREADS, 0, side->st>>>(w.U2, N, Np, w.T, Sp, w.ut);
            CBO_CUDA(cudaEventRecord(side->ev_ut, side->st));
        }
        {
            TimeScope ts_(T_GEMM2, s, 1);
            CBO_CUDA(cbo::launch_gemm_tn(w.V, Np, gp->linv_t, Np, w.UW, Np, n, Np, Np, cbo::TRI_UPPER, nullptr, s, num_sms(), splitk));
        }
        if (side) {   // join: per-train-point tables (probe branch) and the candidates' density ratios ahead of the score kernel
            CBO_CUDA(cudaStreamWaitEvent(s, side->ev_tables, 0));
            CBO_CUDA(cudaStreamWaitEvent(s, side->ev_ut, 0));            // (behind the candidates' density ratios on that stream)
            if (ratio) CBO_CUDA(cudaStreamWaitEvent(s, side->join2, 0));
        }

        // ---- scores / rank / mask / integration ----------------------------------------------------
        cbo::ScoreArgs a;
        memset(&a, 0, sizeof(a));
        a.N = N; a.Np = Np; a.G = G; a.S = (int)p->num_mc; a.Sp = Sp; a.n_cand = ncand; a.tiles_n = tiles_n;
        a.noise = gp->noise; a.kappa = gp->outputscale; a.alpha_lvl = p->alpha; a.temp = p->temp;
        a.beta_prime = p->beta_prime; a.best_f = p->best_f;
        a.k_magic = 6755399441055744.0; a.k_c0 = 0.9999999988534117; a.k_c1n = -0.043321698775061875;
        a.k_c2 = 0.0009384214835622084; a.k_c3n = -1.3551125681305332e-05;
        a.fx_b2 = cbo::kFxB2; a.fx_b1 = cbo::kFxB1; a.fx_b0 = cbo::kFxB0;   // score_rows_fx (soft path of k_score, k_score_ws, k_score_bwd)
        a.unit_dr = ratio ? 0 : 1;
        a.kind = p->kind; a.optimistic = p->optimistic; a.randomized = p->randomized;
        a.eps_per_candidate = p->eps_per_candidate;
        a.a = w.a; a.h0 = w.h0; a.zb = w.zb; a.dr = w.dr_train; a.sum_dr = w.sum_dr; a.z_cand = z + int64_t(N) * Sp;
        a.W = w.UW; a.rn_partial = w.rn; a.mu = w.mu; a.ut = w.ut; a.drc = ratio ? w.drc : nullptr;
        a.eps_grid = p->eps_per_candidate ? eps_grid + c0 * G * q : eps_grid;
        a.eps_mc = eps_mc;
        a.eps_absmax = p->eps_per_candidate ? w.eps_absmax + c0 : w.eps_absmax;
        a.u_rand = u_rand ? u_rand + c0 * G * q : nullptr;
        cbo::ScoreQArgs aq;
        aq.V = w.V; aq.Xq = Xc; aq.inv_ls = gp->inv_lengthscale; aq.d = d;
        aq.post_cov = (q > 1 && out->post_var) ? out->post_var + c0 * q * q : nullptr;
        cbo::NeiArgs na_last;                                     // the joint-sample launch of the final pass (its reverse needs it)
        memset(&na_last, 0, sizeof(na_last));
        auto run_nei = [&](cbo::ScoreArgs& x) -> cudaError_t {
            {
                cbo::NeiArgs na;
                na.N = N; na.Np = Np; na.d = d; na.nb = (int)p->nei->nb; na.G = G; na.S = (int)p->num_mc; na.n_cand = ncand;
                na.tiles_n = tiles_n; na.noise = gp->noise; na.kappa = gp->outputscale; na.temp = p->temp;
                na.Xq = Xc; na.inv_ls = gp->inv_lengthscale; na.Xb = p->nei->Xb; na.KbX = p->nei->KbX; na.Sbb = p->nei->Sbb;
                na.mub = p->nei->mub; na.W = w.UW; na.rn_partial = w.rn; na.mu = w.mu; na.cscale = x.cscale;
                na.eps_grid = x.eps_grid; na.eps_per_candidate = x.eps_per_candidate; na.eps_mc = eps_mc;
                na.values = out->mc_values ? out->mc_values + c0 * q * int64_t(p->num_mc) * G : w.nei_values;
                na.raw = p->kind == CBO_ACQ_JOINT;
                na.V = w.V;
                if (cudaError_t e = cbo::launch_nei_values_q(na, q, s); e != cudaSuccess) return e;
                x.ext_values = na.values;
                na_last = na;
            }
            return cudaGetLastError();
        };
        auto run_score = [&](const cbo::ScoreArgs& x_in) -> cudaError_t {
            cbo::ScoreArgs x = x_in;
            if (nei_kind)
                if (cudaError_t e = run_nei(x); e != cudaSuccess) return e;
            if (q == 1) return soft_path ? launch_score<true>(x, s, many) : launch_score<false>(x, s, many);
            cbo::ScoreQArgs y = aq;
            y.base = x;
            return launch_score_q(y, q, soft_path, s);
        };
        auto set_outputs = [&](cbo::ScoreArgs& x) {
            x.acq = out->acq + c0;
            x.grid = out->target_grid ? out->target_grid + c0 * G * q : nullptr;
            x.logp = out->grid_logp ? out->grid_logp + c0 * G : nullptr;
            x.mask = out->conf_pred_mask ? out->conf_pred_mask + c0 * G * q : nullptr;
            x.qscore = out->q_conf_scores ? out->q_conf_scores + c0 * G * q : nullptr;
            x.ood = out->ood_mask ? out->ood_mask + c0 * q : nullptr;
            x.post_mean = out->post_mean ? out->post_mean + c0 * q : nullptr;
            x.post_var = (q == 1 && out->post_var) ? out->post_var + c0 : nullptr;   // q > 1: full q x q via post_cov
            x.num_acc = out->num_accepted ? out->num_accepted + c0 * q : nullptr;
            x.refit_mean = (q == 1 && out->refit_mean) ? out->refit_mean + c0 * G : nullptr;
            x.refit_sd = (q == 1 && out->refit_sd) ? out->refit_sd + c0 : nullptr;
        };
        // final pass: value only, or value + reverse pass (k_score_bwd -> A^-1 solve of w_bar -> k_rows_bwd)
        auto run_final = [&](const cbo::ScoreArgs& fa) -> int {
            const bool extra = out->target_grid || out->grid_logp || out->conf_pred_mask || out->ood_mask ||
                               out->q_conf_scores || out->post_mean || out->post_var || out->num_accepted ||
                               out->refit_mean || out->refit_sd;
            if (!grad || extra) {
                TimeScope ts_(T_SCORE, s, 1);
                CBO_CUDA(run_score(fa));
            }
            if (!grad) return CBO_OK;
            if (!qpath) {
                TimeScope ts_(T_SCORE_BWD, s, 1);
                cbo::ScoreBwdArgs ba;
                ba.f = fa;
                ba.wbar = w.V;                                       // v is dead after ||v||^2
                ba.adj = w.adj;
                CBO_CUDA(soft_path ? launch_score_bwd<true>(ba, s) : launch_score_bwd<false>(ba, s));
            } else {
                TimeScope ts_(T_SCORE_BWD, s, 1);
                cbo::ScoreQBwdArgs bq;
                memset(&bq, 0, sizeof(bq));
                bq.f = aq;
                bq.f.base = fa;
                bq.f.post_cov = nullptr;
                bq.wbar = w.X3;                                      // v (w.V) is read by phase 0: right-hand sides go to the scratch buffer
                bq.adjq = w.adj;
                bq.grid_leak = p->grid_grad_leak ? 1 : 0;
                if (nei_kind) CBO_CUDA(run_nei(bq.f.base));          // sample values of the joint posterior (constants of k_score_q_bwd)
                if (ext) {
                    bq.ext_mask_bar = ext->mask_bar + c0 * G * q;
                    bq.ext_ood_bar = ext->ood_bar + c0 * q;
                    if (!nei_kind && ext->f_bar) bq.ext_f_bar = ext->f_bar + c0 * int64_t(p->num_mc) * G * q;
                } else if (nei_kind) {
                    bq.val_bar = w.val_bar;
                }
                CBO_CUDA(launch_score_q_bwd(bq, q, soft_path, s));
                if (nei_kind && (!ext || ext->f_bar)) {              // reverse of the joint posterior: adds to adjq / wbar, writes cv_bar
                    cbo::NeiBwdArgs nb_;
                    nb_.f = na_last;
                    nb_.val_bar = ext ? ext->f_bar + c0 * int64_t(p->num_mc) * G * q : w.val_bar;
                    nb_.adjq = w.adj; nb_.wbar = w.X3; nb_.cv_bar = w.cv_bar;
                    nb_.grid_leak = (p->grid_grad_leak && !p->grid_cov_scale && R == 0) ? 1 : 0;
                    CBO_CUDA(cbo::launch_nei_values_q_bwd(nb_, q, s));
                }
            }
            const bool nei_rev = qpath && nei_kind && (!ext || ext->f_bar);
            double* rhs = !qpath ? w.V : w.X3;                      // right-hand sides w_bar'
            double* sol = !qpath ? w.X3 : w.V;                      // scratch of the two-pass solve
            // latency path: the ratio estimator's reverse pass needs only the adjoint records -- it runs on the side stream
            // beside the two reverse solves and WRITES dX, the kernel-row term is then added (same sum, other order of issue)
            const bool ratio_first = side && ratio && (qpath || splitk != nullptr);
            if (ratio_first) {
                CBO_CUDA(cudaEventRecord(side->fork, s));
                CBO_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
                CBO_CUDA(cbo::launch_ratio_mlp_bwd(Xc, n, d, (int)ratio->hidden, ratio->W1, ratio->b1, ratio->W2,
                                                   ratio->b2, ratio->W3, ratio->b3, 1e-6, w.adj,
                                                   !qpath ? cbo::ADJ_STRIDE : cbo::ADJQ_STRIDE, !qpath ? 2 : 1,
                                                   dX + c0 * q * d, num_sms(), side->st, false));
                CBO_CUDA(cudaEventRecord(side->join, side->st));
            }
            {
                TimeScope ts_(T_GEMM_BWD, s, 2);
                CBO_CUDA(cbo::launch_gemm_tn(rhs, Np, gp->linv, Np, sol, Np, n, Np, Np, cbo::TRI_LOWER, nullptr, s, num_sms(), splitk));
                CBO_CUDA(cbo::launch_gemm_tn(sol, Np, gp->linv_t, Np, rhs, Np, n, Np, Np, cbo::TRI_UPPER, nullptr, s, num_sms(), splitk));
            }
            {
                TimeScope ts_(T_ROWS_BWD, s, ratio ? 2 : 1);
                if (!qpath && splitk == nullptr)
                    CBO_CUDA(cbo::launch_rows_bwd(Xc, n, w.XsT, N, Np, d, gp->inv_lengthscale, gp->outputscale, gp->alpha, rhs,
                                                  w.UW, w.adj, dX + c0 * q * d, num_sms(), s));
                else {  // joint batches, and q = 1 calls with a handful of candidates (CTA per candidate point)
                    if (ratio_first) CBO_CUDA(cudaStreamWaitEvent(s, side->join, 0));
                    if (splitk != nullptr && N > 512)   // (per call: a handful of rows, enough train points for 32 warps)
                        CBO_LAUNCHED(), cbo::k_rows_bwd_q<1024><<<n, 1024, 0, s>>>(Xc, n, q, w.XsT, N, Np, d, gp->inv_lengthscale, gp->outputscale,
                                                                 gp->alpha, rhs, w.UW, w.adj,
                                                                 !qpath ? cbo::ADJ_STRIDE : cbo::ADJQ_STRIDE, !qpath ? 1 : 0,
                                                                 nei_rev ? p->nei->Xb : nullptr, nei_rev ? (int)p->nei->nb : 0,
                                                                 nei_rev ? w.cv_bar : nullptr, dX + c0 * q * d, ratio_first ? 1 : 0);
                    else
                        CBO_LAUNCHED(), cbo::k_rows_bwd_q<256><<<n, 256, 0, s>>>(Xc, n, q, w.XsT, N, Np, d, gp->inv_lengthscale, gp->outputscale,
                                                                 gp->alpha, rhs, w.UW, w.adj,
                                                                 !qpath ? cbo::ADJ_STRIDE : cbo::ADJQ_STRIDE, !qpath ? 1 : 0,
                                                                 nei_rev ? p->nei->Xb : nullptr, nei_rev ? (int)p->nei->nb : 0,
                                                                 nei_rev ? w.cv_bar : nullptr, dX + c0 * q * d, ratio_first ? 1 : 0);
                }
                if (ratio && !ratio_first)
                    CBO_CUDA(cbo::launch_ratio_mlp_bwd(Xc, n, d, (int)ratio->hidden, ratio->W1, ratio->b1, ratio->W2,
                                                       ratio->b2, ratio->W3, ratio->b3, 1e-6, w.adj,
                                                       !qpath ? cbo::ADJ_STRIDE : cbo::ADJQ_STRIDE, !qpath ? 2 : 1,
                                                       dX + c0 * q * d, num_sms(), s));
                CBO_CUDA(cudaGetLastError());
            }
            return CBO_OK;
        };
        if (R == 0) {
            if (p->grid_cov_scale) a.cscale = p->grid_cov_scale + c0 * q;
            set_outputs(a);
            int rs = run_final(a);
            if (rs != CBO_OK) return rs;
        } else {
            // grid refinement (helpers.py:129-189): probe passes keep only the accepted counts; the pass with the
            // smallest grid_score is then re-run into the caller's buffers.  One host sync per pass, like the
            // reference (helpers.py:151-152,177).
            if (p->grid_cov_scale) CBO_CUDA(cudaMemcpyAsync(w.cscale, p->grid_cov_scale, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
            else CBO_LAUNCHED(), cbo::k_fill<<<(n + 255) / 256, 256, 0, s>>>(w.cscale, n, 1.0);
            float best_score = INFINITY;
            int best_step = 0;
            double* scratch_acq = w.acq_scratch;          // acquisition values of probe passes are discarded
            for (int step = 0; step <= R; ++step) {
                cbo::ScoreArgs pa = a;
                pa.cscale = w.cscale + int64_t(step) * b * q;
                pa.acq = scratch_acq;
                pa.num_acc = w.nacc;
                CBO_CUDA(run_score(pa));
                CBO_LAUNCHED(), k_accept_stats<<<1, 256, 0, s>>>(w.nacc, n, G, w.stats);
                double hs[4];
                CBO_CUDA(cudaMemcpyAsync(hs, w.stats, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
                CBO_CUDA(cudaStreamSynchronize(s));
                hs[3] = double(n);
                // candidate-sharded calls: the reference's decisions are batch-global (helpers.py:151-160,177) -- the caller's
                // hook reduces {min, max, sum, count} over the ranks (one tiny all-reduce per refinement step)
                if (p->refine_reduce) p->refine_reduce(hs, p->refine_ctx);
                float score = (float)hs[2] / (float)hs[3];
                if (score < best_score) { best_score = score; best_step = step; }
                bool rescale = hs[0] < 2.0 || hs[1] > double(G - 2);
                if (!rescale) break;
                if (step < R)
                    CBO_LAUNCHED(), k_rescale<<<(n + 255) / 256, 256, 0, s>>>(w.nacc, w.cscale + int64_t(step) * b * q,
                                                              w.cscale + int64_t(step + 1) * b * q, n, G, q);
            }
            cbo::ScoreArgs fa = a;
            fa.cscale = w.cscale + int64_t(best_step) * b * q;
            set_outputs(fa);
            int rs = run_final(fa);
            if (rs != CBO_OK) return rs;
        }
        if (hooks) { int hs_ = hooks->after(hooks->ctx, c0, ncand); if (hs_ != CBO_OK) return hs_; }
    }
    return CBO_OK;
}

int cbo_conformal_acq_f64(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* p, const double* Xq,
                          int64_t b, const double* eps_grid, const double* z, const double* eps_mc,
                          const double* u_rand, const cbo_outputs_f64* out, void* workspace, size_t workspace_bytes,
                          int64_t chunk, void* stream) {
    return conformal_acq_impl(gp, ratio, p, Xq, b, eps_grid, z, eps_mc, u_rand, out, nullptr, workspace, workspace_bytes,
                              chunk, stream);
}

int cbo_conformal_acq_grad_f64(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* p, const double* Xq,
                               int64_t b, const double* eps_grid, const double* z, const double* eps_mc,
                               const double* u_rand, const cbo_outputs_f64* out, double* dacq_dX, void* workspace,
                               size_t workspace_bytes, int64_t chunk, void* stream) {
    if (!dacq_dX) return CBO_ERR_INVALID_ARG;
    return conformal_acq_impl(gp, ratio, p, Xq, b, eps_grid, z, eps_mc, u_rand, out, dacq_dX, workspace, workspace_bytes,
                              chunk, stream);
}

size_t cbo_workspace_bytes_grad(int64_t N, int64_t d, int64_t b, const cbo_params* params, int64_t chunk) {
    if (!params || N <= 0 || b <= 0) return 0;
    return carve(nullptr, N, d, b, params, chunk, true).bytes;
}

int cbo_cehvi_integrate_f64(int64_t m, int64_t b, int64_t G, int64_t S, const double* const* mask,
                            const double* const* logp, const double* const* grid, const double* const* mu,
                            const double* const* ood, const double* const* refit_mean, const double* const* refit_sd,
                            const double* eps_mc, const double* cell_lower, const double* cell_upper, int64_t n_cells,
                            double alpha, int32_t optimistic, double* acq, void* stream) {
    if (!mask || !logp || !grid || !mu || !ood || !refit_mean || !refit_sd || !eps_mc || !cell_lower || !cell_upper || !acq)
        return CBO_ERR_INVALID_ARG;
    if (m < 1 || b <= 0 || G < 2 || S < 1 || n_cells < 1) return CBO_ERR_INVALID_ARG;
    if (m > cbo::MAX_M || size_t(2 * n_cells * m) * sizeof(double) > 200 * 1024) return CBO_ERR_UNSUPPORTED;
    cbo::EhviArgs a;
    memset(&a, 0, sizeof(a));
    a.m = (int)m; a.b = (int)b; a.G = (int)G; a.S = (int)S; a.n_cells = (int)n_cells; a.optimistic = optimistic;
    a.alpha_lvl = alpha;
    for (int k = 0; k < m; ++k) {
        if (!mask[k] || !logp[k] || !grid[k] || !mu[k] || !ood[k] || !refit_mean[k] || !refit_sd[k]) return CBO_ERR_INVALID_ARG;
        a.mask[k] = mask[k]; a.logp[k] = logp[k]; a.grid[k] = grid[k]; a.mu[k] = mu[k]; a.ood[k] = ood[k];
        a.mf[k] = refit_mean[k]; a.sd[k] = refit_sd[k];
    }
    a.eps_mc = eps_mc; a.cell_lower = cell_lower; a.cell_upper = cell_upper; a.acq = acq;
    size_t smem = size_t(2 * n_cells * m) * sizeof(double);
    static cbo::SmemOptIn optin;
    CBO_CUDA(optin.ensure(cbo::k_ehvi, smem));
    CBO_LAUNCHED(), cbo::k_ehvi<<<(int)((b + 7) / 8), 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    CBO_CUDA(cudaGetLastError());
    return CBO_OK;
}

int cbo_cehvi_integrate_q_f64(int64_t m, int64_t b, int64_t G, int64_t S, int64_t q, const double* const* mask,
                              const double* const* logp, const double* const* grid, const double* const* mu,
                              const double* const* ood, const double* const* cov, const double* noise_host,
                              const double* const* fx, const double* eps_mc, const double* cell_lower,
                              const double* cell_upper, int64_t n_cells, int32_t cells_per_sample, double alpha,
                              int32_t optimistic, double* acq, double* values_out, void* stream) {
    return cehvi_integrate_q_impl(m, b, G, S, q, mask, logp, grid, mu, ood, cov, noise_host, fx, eps_mc, cell_lower, cell_upper,
                                  n_cells, cells_per_sample, alpha, optimistic, acq, values_out, false, nullptr, nullptr,
                                  nullptr, stream);
}

int cbo_cehvi_integrate_q_grad_f64(int64_t m, int64_t b, int64_t G, int64_t S, int64_t q, const double* const* mask,
                                   const double* const* logp, const double* const* grid, const double* const* mu,
                                   const double* const* ood, const double* const* cov, const double* noise_host,
                                   const double* const* fx, const double* eps_mc, const double* cell_lower,
                                   const double* cell_upper, int64_t n_cells, int32_t cells_per_sample, double alpha,
                                   int32_t optimistic, double* acq, double* values_out, double* const* mask_bar,
                                   double* const* ood_bar, double* const* f_bar, void* stream) {
    return cehvi_integrate_q_impl(m, b, G, S, q, mask, logp, grid, mu, ood, cov, noise_host, fx, eps_mc, cell_lower, cell_upper,
                                  n_cells, cells_per_sample, alpha, optimistic, acq, values_out, true, mask_bar, ood_bar,
                                  f_bar, stream);
}

size_t cbo_model_workspace_bytes(int64_t N, int64_t d) {
    if (N <= 0 || d <= 0) return 0;
    return cbo::model_build_carve(nullptr, N, d).bytes;
}

int cbo_model_build_f64(const double* X, const double* y, int64_t N, int64_t d, const double* inv_lengthscale, double mean,
                        double outputscale, double noise, double* alpha, double* linv, double* linv_t, int32_t* info,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!X || !y || !inv_lengthscale || !alpha || !linv || !linv_t || !workspace || N <= 0 || d <= 0) return CBO_ERR_INVALID_ARG;
    if (d > cbo::MAX_D || N > 32768) return CBO_ERR_UNSUPPORTED;
    if (!(outputscale > 0.0) || !(noise > 0.0)) return CBO_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return CBO_ERR_WORKSPACE;
    cbo::ModelBuildWs w = cbo::model_build_carve(workspace, N, d);
    if (w.bytes > workspace_bytes) return CBO_ERR_WORKSPACE;
    CBO_CUDA(cbo::model_build(X, y, (int)N, (int)d, inv_lengthscale, mean, outputscale, noise, alpha, linv, linv_t, info, w,
                              num_sms(), static_cast<cudaStream_t>(stream)));
    return CBO_OK;
}

#ifdef CBO_WS_PROFILE
int cbo_debug_ws_prof(unsigned long long* out8) {   // experiment builds only (tools/score_ab.py): read and clear the counters
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out8, cbo::g_ws_prof, sizeof(z)) != cudaSuccess) return CBO_ERR_CUDA;
    if (cudaMemcpyToSymbol(cbo::g_ws_prof, z, sizeof(z)) != cudaSuccess) return CBO_ERR_CUDA;
    return CBO_OK;
}
#endif

void cbo_timing_enable(int on) {
    g_timing.enabled = on != 0;
}

// Sums elapsed ms per kernel class since the last collect (synchronises on the recorded events).
int cbo_timing_collect(double* ms_out /* 8 */, int64_t* launches_out /* 8 */) {
    if (!ms_out || !launches_out) return CBO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> l(g_timing.mu);
    for (int c = 0; c < T_NCLASS; ++c) {
        double tot = 0.0;
        auto& v = g_timing.ev[c];
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            CBO_CUDA(cudaEventSynchronize(v[i + 1]));
            float ms = 0.f;
            CBO_CUDA(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
            tot += ms;
        }
        for (cudaEvent_t e : v) g_timing.pool.push_back(e);
        v.clear();
        ms_out[c] = tot;
        launches_out[c] = g_timing.launches[c];
        g_timing.launches[c] = 0;
    }
    return CBO_OK;
}

}  // extern "C"

// ---- host-buffer entry point: per-chunk pipeline ------------------------------------------------------------------
// The candidates of chunk k+1 travel host -> device and the acquisition values of chunk k-1 device -> host on a private copy
// stream while the kernels of chunk k run on the caller's stream (events order the two); one copy stream and a pool of
// timing-free events per device, created on first use (no device memory is allocated).
namespace {
struct HostPipe {
    std::mutex mu;
    cudaStream_t copy = nullptr;
    std::vector<cudaEvent_t> ev;
    cudaError_t ensure(size_t n_events) {
        if (!copy) {
            cudaError_t e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        while (ev.size() < n_events) {
            cudaEvent_t x;
            cudaError_t e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            ev.push_back(x);
        }
        return cudaSuccess;
    }
};
HostPipe g_host_pipe[cbo::kMaxDevices];

struct HostPipeCtx {
    HostPipe* pipe;
    cudaStream_t compute;
    const double* X_host; double* X_dev; double* acq_host; double* acq_dev;
    int64_t b, C, row;          // candidates, chunk size, doubles per candidate (q * d)
    cudaError_t err = cudaSuccess;
    cudaEvent_t h2d_done(int64_t k) { return pipe->ev[2 * k + 1]; }
    cudaEvent_t chunk_done(int64_t k) { return pipe->ev[2 * k + 2]; }
    cudaError_t upload(int64_t k) {     // candidates of chunk k
        const int64_t c0 = k * C, n = (b - c0 < C) ? (b - c0) : C;
        cudaError_t e = cudaMemcpyAsync(X_dev + c0 * row, X_host + c0 * row, sizeof(double) * n * row, cudaMemcpyHostToDevice, pipe->copy);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(h2d_done(k), pipe->copy);
    }
};
int host_before(void* v, int64_t c0, int64_t) {
    HostPipeCtx* c = static_cast<HostPipeCtx*>(v);
    const int64_t k = c0 / c->C;
    if ((k + 1) * c->C < c->b) { if ((c->err = c->upload(k + 1)) != cudaSuccess) return CBO_ERR_CUDA; }   // prefetch the next chunk
    if ((c->err = cudaStreamWaitEvent(c->compute, c->h2d_done(k), 0)) != cudaSuccess) return CBO_ERR_CUDA;
    return CBO_OK;
}
int host_after(void* v, int64_t c0, int64_t ncand) {
    HostPipeCtx* c = static_cast<HostPipeCtx*>(v);
    const int64_t k = c0 / c->C;
    if ((c->err = cudaEventRecord(c->chunk_done(k), c->compute)) != cudaSuccess) return CBO_ERR_CUDA;
    if ((c->err = cudaStreamWaitEvent(c->pipe->copy, c->chunk_done(k), 0)) != cudaSuccess) return CBO_ERR_CUDA;
    if ((c->err = cudaMemcpyAsync(c->acq_host + c0, c->acq_dev + c0, sizeof(double) * ncand, cudaMemcpyDeviceToHost,
                                  c->pipe->copy)) != cudaSuccess) return CBO_ERR_CUDA;
    return CBO_OK;
}
}  // namespace

extern "C" int cbo_conformal_acq_host_f64(const cbo_gp_f64* gp, const cbo_ratio_mlp_f64* ratio, const cbo_params* p,
                               const double* Xq_host, int64_t b, const double* eps_grid, const double* z,
                               const double* eps_mc, const double* u_rand, double* acq_host, double* Xq_staging,
                               double* acq_staging, void* workspace, size_t workspace_bytes, int64_t chunk,
                               void* stream) {
    if (!gp || !p || !Xq_host || !acq_host || !Xq_staging || !acq_staging || b <= 0) return CBO_ERR_INVALID_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int dev = cbo::current_device_slot();
    if (dev < 0) return CBO_ERR_NO_DEVICE;
    HostPipe& pipe = g_host_pipe[dev];
    std::lock_guard<std::mutex> lock(pipe.mu);            // one host-buffer call per device at a time shares the copy stream
    const int64_t C = pick_chunk(b, p, chunk), n_chunks = (b + C - 1) / C;
    CBO_CUDA(pipe.ensure(size_t(2 * n_chunks + 2)));
    HostPipeCtx ctx;
    ctx.pipe = &pipe; ctx.compute = s; ctx.X_host = Xq_host; ctx.X_dev = Xq_staging; ctx.acq_host = acq_host;
    ctx.acq_dev = acq_staging; ctx.b = b; ctx.C = C; ctx.row = p->q * gp->d;
    // the copy stream starts behind whatever the caller's stream still has in flight on the staging buffers
    CBO_CUDA(cudaEventRecord(pipe.ev[0], s));
    CBO_CUDA(cudaStreamWaitEvent(pipe.copy, pipe.ev[0], 0));
    CBO_CUDA(ctx.upload(0));
    cbo_outputs_f64 out;
    memset(&out, 0, sizeof(out));
    out.acq = acq_staging;
    ChunkHooks hooks{&ctx, host_before, host_after};
    int st = conformal_acq_impl(gp, ratio, p, Xq_staging, b, eps_grid, z, eps_mc, u_rand, &out, nullptr, workspace,
                                workspace_bytes, chunk, stream, &hooks);
    if (st == CBO_ERR_CUDA && ctx.err != cudaSuccess) cuda_fail(ctx.err, "host pipeline");
    // drain both streams before returning (the host buffers belong to the caller again)
    cudaError_t e1 = cudaStreamSynchronize(pipe.copy), e2 = cudaStreamSynchronize(s);
    if (st != CBO_OK) return st;
    CBO_CUDA(e1);
    CBO_CUDA(e2);
    return CBO_OK;
}
