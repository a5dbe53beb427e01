// torch_ext.cpp -- thin torch C++ extension over the C ABI of libcbo (include/cbo.h): TORCH_LIBRARY(cbo, ...).
//
// The operators take and return torch tensors, allocate outputs and workspace through torch's caching allocator, enqueue on
// torch's current CUDA stream and turn non-zero cbo_status codes into TORCH_CHECK failures ("the torch
// extension converts non-zero to TORCH_CHECK failures").  No arithmetic lives here -- every operator is one or two calls
// into libcbo.so.  `cbo::conformal_acq` is differentiable in the candidates: its Autograd kernel runs
// cbo_conformal_acq_grad_f64 once and hands the stored Jacobian rows to backward, which is what the reference gets from
// autograd through conformalbo/acquisitions.py:53-76 inside optimize_acqf_sgld (conformalbo/optim/optimize.py:224-235).
//
//   torch.ops.cbo.conformal_acq        ConformalAcquisition.forward for cUCB / cEI (acquisitions.py:73-76, 88-98, 151-159)
//   torch.ops.cbo.conformal_acq_raw    the same without autograd: (acq, d acq / d X or an empty tensor)
//   torch.ops.cbo.model_build          model refit -> alpha, L^-1, L^-T (conformalbo/utils.py:382-390)
//   torch.ops.cbo.conf_mask_to_bounds  helpers.py:15-91
//   torch.ops.cbo.ratio_forward        RatioEstimator.forward (ratio_estimation.py:97-99)
//   torch.ops.cbo.ratio_train_step     RatioEstimator.optimize_callback from the minibatch on (ratio_estimation.py:133-144)
//   torch.ops.cbo.sgld_update          one SGLD / SGHMC step (optim/sgld.py:9-47, optim/optimize.py:224-235)
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/torch.h>

#include <cstring>
#include <tuple>
#include <vector>

#include "../../include/cbo.h"

namespace {

using torch::Tensor;

void check_status(int st, const char* what) {
    TORCH_CHECK(st == CBO_OK, what, ": ", cbo_status_str(st), (st == CBO_ERR_CUDA ? std::string(" (") + cbo_last_cuda_error() + ")" : std::string()));
}

Tensor dev_f64(const Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (libcbo has no CPU path)");
    return t.to(torch::kFloat64).contiguous();
}

void* stream_ptr() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

Tensor aligned_workspace(size_t bytes, const Tensor& like, char** ptr) {
    Tensor ws = torch::empty({static_cast<int64_t>(bytes + 256)}, like.options().dtype(torch::kUInt8));
    uintptr_t p = reinterpret_cast<uintptr_t>(ws.data_ptr());
    *ptr = reinterpret_cast<char*>((p + 255) / 256 * 256);
    return ws;
}

struct Gp {
    Tensor X, alpha, linv, linv_t, inv_ls;
    cbo_gp_f64 c;
};

Gp make_gp(const Tensor& X, const Tensor& alpha, const Tensor& linv, const Tensor& linv_t, const Tensor& inv_ls, double mean,
           double outputscale, double noise) {
    Gp g;
    g.X = dev_f64(X, "X_train"); g.alpha = dev_f64(alpha, "alpha"); g.linv = dev_f64(linv, "linv");
    g.linv_t = dev_f64(linv_t, "linv_t"); g.inv_ls = dev_f64(inv_ls, "inv_lengthscale");
    TORCH_CHECK(g.X.dim() == 2, "X_train must be N x d");
    std::memset(&g.c, 0, sizeof(g.c));
    g.c.N = g.X.size(0); g.c.d = g.X.size(1); g.c.Np = cbo_padded_n(g.c.N);
    TORCH_CHECK(g.alpha.numel() == g.c.Np && g.linv.numel() == g.c.Np * g.c.Np && g.linv_t.numel() == g.c.Np * g.c.Np,
                "alpha / linv / linv_t must be padded to Np = ", g.c.Np, " (cbo_padded_n)");
    TORCH_CHECK(g.inv_ls.numel() == g.c.d, "inv_lengthscale must hold d values");
    g.c.X = g.X.data_ptr<double>(); g.c.alpha = g.alpha.data_ptr<double>(); g.c.linv = g.linv.data_ptr<double>();
    g.c.linv_t = g.linv_t.data_ptr<double>(); g.c.inv_lengthscale = g.inv_ls.data_ptr<double>();
    g.c.mean = mean; g.c.outputscale = outputscale; g.c.noise = noise;
    return g;
}

struct Ratio {
    std::vector<Tensor> w;
    cbo_ratio_mlp_f64 c;
    bool present = false;
};

Ratio make_ratio(const std::vector<Tensor>& weights) {
    Ratio r;
    if (weights.empty()) return r;
    TORCH_CHECK(weights.size() == 6, "ratio weights: [W1, b1, W2, b2, W3, b3]");
    for (const Tensor& t : weights) r.w.push_back(dev_f64(t, "ratio weight"));
    r.c.hidden = r.w[1].numel();
    r.c.W1 = r.w[0].data_ptr<double>(); r.c.b1 = r.w[1].data_ptr<double>(); r.c.W2 = r.w[2].data_ptr<double>();
    r.c.b2 = r.w[3].data_ptr<double>(); r.c.W3 = r.w[4].data_ptr<double>(); r.c.b3 = r.w[5].data_ptr<double>();
    r.present = true;
    return r;
}

// (acq (b), dacq_dX (b x q x d) or an empty tensor)
std::tuple<Tensor, Tensor> conformal_acq_raw(const Tensor& X_train, const Tensor& alpha, const Tensor& linv, const Tensor& linv_t,
                                             const Tensor& inv_ls, double mean, double outputscale, double noise, const Tensor& Xq,
                                             const Tensor& eps_grid, const Tensor& z, const Tensor& eps_mc,
                                             const c10::optional<Tensor>& u_rand, std::vector<Tensor> ratio, int64_t kind,
                                             double alpha_lvl, double temp, int64_t grid_res, int64_t max_grid_refinements,
                                             bool optimistic, bool randomized, double beta_prime, double best_f, int64_t chunk,
                                             bool want_grad) {
    c10::cuda::CUDAGuard guard(Xq.device());
    Gp gp = make_gp(X_train, alpha, linv, linv_t, inv_ls, mean, outputscale, noise);
    Ratio rt = make_ratio(ratio);
    Tensor xq = dev_f64(Xq.dim() == 2 ? Xq.unsqueeze(1) : Xq, "Xq");
    TORCH_CHECK(xq.dim() == 3 && xq.size(2) == gp.c.d, "Xq must be b x q x d");
    TORCH_CHECK(kind == CBO_ACQ_UCB || kind == CBO_ACQ_EI, "torch.ops.cbo.conformal_acq covers cUCB (0) and cEI (1)");
    const int64_t b = xq.size(0), q = xq.size(1), G = grid_res;
    Tensor eg = dev_f64(eps_grid, "eps_grid"), zz = dev_f64(z, "z"), em = dev_f64(eps_mc, "eps_mc");
    TORCH_CHECK(eg.numel() == G * q || eg.numel() == b * G * q, "eps_grid must hold G*q or b*G*q values");
    TORCH_CHECK(zz.dim() == 2 && zz.size(0) == gp.c.N + q, "z must be (N + q) x S_p");
    TORCH_CHECK(em.numel() % (G * q) == 0 && em.numel() > 0, "eps_mc must be S x G x q");
    Tensor ur;
    if (u_rand.has_value() && u_rand->defined()) {
        ur = dev_f64(*u_rand, "u_rand");
        TORCH_CHECK(ur.numel() == b * G * q, "u_rand must hold b*G*q uniforms");
    }
    TORCH_CHECK(!randomized || ur.defined(), "randomized=True needs u_rand");
    cbo_params p;
    std::memset(&p, 0, sizeof(p));
    p.alpha = alpha_lvl; p.temp = temp; p.grid_res = G; p.q = q; p.num_mc = em.numel() / (G * q); p.num_probes = zz.size(1);
    p.max_grid_refinements = max_grid_refinements; p.kind = static_cast<int32_t>(kind); p.optimistic = optimistic;
    p.randomized = randomized; p.eps_per_candidate = (eg.numel() == b * G * q && eg.numel() != G * q);
    p.beta_prime = beta_prime; p.best_f = best_f;
    Tensor acq = torch::empty({b}, xq.options());
    Tensor dX = want_grad ? torch::empty({b, q, gp.c.d}, xq.options()) : torch::empty({0}, xq.options());
    cbo_outputs_f64 out;
    std::memset(&out, 0, sizeof(out));
    out.acq = acq.data_ptr<double>();
    const size_t bytes = want_grad ? cbo_workspace_bytes_grad(gp.c.N, gp.c.d, b, &p, chunk) : cbo_workspace_bytes(gp.c.N, gp.c.d, b, &p, chunk);
    char* wp = nullptr;
    Tensor ws = aligned_workspace(bytes, xq, &wp);
    const double* urp = ur.defined() ? ur.data_ptr<double>() : nullptr;
    int st;
    if (want_grad)
        st = cbo_conformal_acq_grad_f64(&gp.c, rt.present ? &rt.c : nullptr, &p, xq.data_ptr<double>(), b, eg.data_ptr<double>(),
                                        zz.data_ptr<double>(), em.data_ptr<double>(), urp, &out, dX.data_ptr<double>(), wp, bytes,
                                        chunk, stream_ptr());
    else
        st = cbo_conformal_acq_f64(&gp.c, rt.present ? &rt.c : nullptr, &p, xq.data_ptr<double>(), b, eg.data_ptr<double>(),
                                   zz.data_ptr<double>(), em.data_ptr<double>(), urp, &out, wp, bytes, chunk, stream_ptr());
    check_status(st, want_grad ? "cbo_conformal_acq_grad_f64" : "cbo_conformal_acq_f64");
    return {acq, dX};
}

class ConformalAcqFn : public torch::autograd::Function<ConformalAcqFn> {
   public:
    static Tensor forward(torch::autograd::AutogradContext* ctx, const Tensor& Xq, const Tensor& X_train, const Tensor& alpha,
                          const Tensor& linv, const Tensor& linv_t, const Tensor& inv_ls, double mean, double outputscale,
                          double noise, const Tensor& eps_grid, const Tensor& z, const Tensor& eps_mc,
                          const c10::optional<Tensor>& u_rand, std::vector<Tensor> ratio, int64_t kind, double alpha_lvl, double temp,
                          int64_t grid_res, int64_t max_grid_refinements, bool optimistic, bool randomized, double beta_prime,
                          double best_f, int64_t chunk) {
        at::AutoDispatchBelowADInplaceOrView guard;
        const bool want_grad = Xq.requires_grad();
        auto [acq, dX] = conformal_acq_raw(X_train, alpha, linv, linv_t, inv_ls, mean, outputscale, noise, Xq, eps_grid, z, eps_mc,
                                           u_rand, std::move(ratio), kind, alpha_lvl, temp, grid_res, max_grid_refinements,
                                           optimistic, randomized, beta_prime, best_f, chunk, want_grad);
        if (want_grad) ctx->save_for_backward({dX});
        ctx->saved_data["two_d"] = Xq.dim() == 2;
        return acq;
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx, torch::autograd::variable_list grad_out) {
        torch::autograd::variable_list g(24);
        auto saved = ctx->get_saved_variables();
        if (!saved.empty()) {
            Tensor dX = saved[0];                                             // acq[c] depends on Xq[c] only: a row scaling
            Tensor gx = grad_out[0].to(dX.dtype()).reshape({-1, 1, 1}) * dX;
            g[0] = ctx->saved_data["two_d"].toBool() ? gx.squeeze(1) : gx;
        }
        return g;
    }
};

Tensor conformal_acq_autograd(const Tensor& X_train, const Tensor& alpha, const Tensor& linv, const Tensor& linv_t, const Tensor& inv_ls,
                              double mean, double outputscale, double noise, const Tensor& Xq, const Tensor& eps_grid, const Tensor& z,
                              const Tensor& eps_mc, const c10::optional<Tensor>& u_rand, std::vector<Tensor> ratio, int64_t kind,
                              double alpha_lvl, double temp, int64_t grid_res, int64_t max_grid_refinements, bool optimistic,
                              bool randomized, double beta_prime, double best_f, int64_t chunk) {
    return ConformalAcqFn::apply(Xq, X_train, alpha, linv, linv_t, inv_ls, mean, outputscale, noise, eps_grid, z, eps_mc, u_rand,
                                 std::move(ratio), kind, alpha_lvl, temp, grid_res, max_grid_refinements, optimistic, randomized,
                                 beta_prime, best_f, chunk);
}

Tensor conformal_acq_cuda(const Tensor& X_train, const Tensor& alpha, const Tensor& linv, const Tensor& linv_t, const Tensor& inv_ls,
                          double mean, double outputscale, double noise, const Tensor& Xq, const Tensor& eps_grid, const Tensor& z,
                          const Tensor& eps_mc, const c10::optional<Tensor>& u_rand, std::vector<Tensor> ratio, int64_t kind,
                          double alpha_lvl, double temp, int64_t grid_res, int64_t max_grid_refinements, bool optimistic,
                          bool randomized, double beta_prime, double best_f, int64_t chunk) {
    return std::get<0>(conformal_acq_raw(X_train, alpha, linv, linv_t, inv_ls, mean, outputscale, noise, Xq, eps_grid, z, eps_mc, u_rand,
                                         std::move(ratio), kind, alpha_lvl, temp, grid_res, max_grid_refinements, optimistic,
                                         randomized, beta_prime, best_f, chunk, false));
}

// raw training data + fixed hyper-parameters -> (alpha (Np), linv (Np x Np), linv_t (Np x Np), info (1, int32))
std::tuple<Tensor, Tensor, Tensor, Tensor> model_build(const Tensor& X, const Tensor& y, const Tensor& inv_ls, double mean,
                                                       double outputscale, double noise) {
    c10::cuda::CUDAGuard guard(X.device());
    Tensor x = dev_f64(X, "X"), yy = dev_f64(y.reshape({-1}), "y"), il = dev_f64(inv_ls, "inv_lengthscale");
    TORCH_CHECK(x.dim() == 2 && yy.numel() == x.size(0) && il.numel() == x.size(1), "X: N x d, y: N, inv_lengthscale: d");
    const int64_t N = x.size(0), d = x.size(1), Np = cbo_padded_n(N);
    Tensor alpha = torch::empty({Np}, x.options()), linv = torch::empty({Np, Np}, x.options()), linv_t = torch::empty({Np, Np}, x.options());
    Tensor info = torch::zeros({1}, x.options().dtype(torch::kInt32));
    const size_t bytes = cbo_model_workspace_bytes(N, d);
    char* wp = nullptr;
    Tensor ws = aligned_workspace(bytes, x, &wp);
    check_status(cbo_model_build_f64(x.data_ptr<double>(), yy.data_ptr<double>(), N, d, il.data_ptr<double>(), mean, outputscale, noise,
                                     alpha.data_ptr<double>(), linv.data_ptr<double>(), linv_t.data_ptr<double>(),
                                     info.data_ptr<int32_t>(), wp, bytes, stream_ptr()),
                 "cbo_model_build_f64");
    return {alpha, linv, linv_t, info};
}

// grid, mask: b x G x 1 x t (or b x G x t) -> lb, ub: b x 1 x t
std::tuple<Tensor, Tensor> conf_mask_to_bounds(const Tensor& target_grid, const Tensor& conf_pred_mask) {
    c10::cuda::CUDAGuard guard(target_grid.device());
    TORCH_CHECK(target_grid.sizes() == conf_pred_mask.sizes(), "grid and mask must have the same shape");
    TORCH_CHECK(target_grid.dim() == 3 || (target_grid.dim() == 4 && target_grid.size(2) == 1), "q = 1 only (helpers.py:35-37)");
    const int64_t b = target_grid.size(0), G = target_grid.size(1), t = target_grid.size(-1);
    Tensor g = dev_f64(target_grid.reshape({b, G, t}), "target_grid"), m = dev_f64(conf_pred_mask.reshape({b, G, t}), "conf_pred_mask");
    Tensor lb = torch::empty({b, t}, g.options()), ub = torch::empty({b, t}, g.options());
    check_status(cbo_conf_mask_to_bounds_f64(g.data_ptr<double>(), m.data_ptr<double>(), b, G, t, lb.data_ptr<double>(),
                                             ub.data_ptr<double>(), stream_ptr()),
                 "cbo_conf_mask_to_bounds_f64");
    return {lb.unsqueeze(1), ub.unsqueeze(1)};
}

Tensor ratio_forward(std::vector<Tensor> weights, const Tensor& x) {
    c10::cuda::CUDAGuard guard(x.device());
    Ratio r = make_ratio(weights);
    TORCH_CHECK(r.present, "ratio weights: [W1, b1, W2, b2, W3, b3]");
    Tensor xx = dev_f64(x, "x");
    const int64_t d = xx.size(-1), n = xx.numel() / d;
    Tensor out = torch::empty({n}, xx.options());
    check_status(cbo_ratio_forward_f64(&r.c, xx.data_ptr<double>(), n, d, out.data_ptr<double>(), stream_ptr()), "cbo_ratio_forward_f64");
    std::vector<int64_t> shape(xx.sizes().begin(), xx.sizes().end() - 1);
    return out.reshape(shape);
}

// in place on theta / adam_m / adam_v / target (flat fp64 vectors in torch's parameter order); returns the minibatch loss (1)
Tensor ratio_train_step(Tensor theta, Tensor adam_m, Tensor adam_v, Tensor target, const Tensor& x_neg, const Tensor& x_pos,
                        const c10::optional<Tensor>& pos_idx, int64_t n_pos, int64_t step, double lr, double beta1, double beta2,
                        double eps, double weight_decay, double ema_weight) {
    c10::cuda::CUDAGuard guard(theta.device());
    for (const Tensor* t : {&theta, &adam_m, &adam_v, &target})
        TORCH_CHECK(t->is_cuda() && t->scalar_type() == torch::kFloat64 && t->is_contiguous(), "parameter vectors must be contiguous fp64 CUDA tensors (updated in place)");
    Tensor xn = dev_f64(x_neg, "x_neg"), xp = dev_f64(x_pos, "x_pos");
    const int64_t d = xp.size(-1), n_neg = xn.numel() / d;
    Tensor idx;
    if (pos_idx.has_value() && pos_idx->defined()) {
        TORCH_CHECK(pos_idx->is_cuda(), "pos_idx must be a CUDA tensor");
        idx = pos_idx->to(torch::kInt64).contiguous();
    }
    Tensor loss = torch::empty({1}, theta.options());
    const size_t bytes = cbo_ratio_train_workspace_bytes(n_neg + n_pos, d);
    TORCH_CHECK(bytes > 0, "unsupported shape (d <= 64, at least one row)");
    char* wp = nullptr;
    Tensor ws = aligned_workspace(bytes, theta, &wp);
    check_status(cbo_ratio_train_step_f64(d, 64, theta.data_ptr<double>(), adam_m.data_ptr<double>(), adam_v.data_ptr<double>(),
                                          target.data_ptr<double>(), xn.data_ptr<double>(), n_neg, xp.data_ptr<double>(),
                                          idx.defined() ? idx.data_ptr<int64_t>() : nullptr, n_pos, step, lr, beta1, beta2, eps,
                                          weight_decay, ema_weight, loss.data_ptr<double>(), wp, bytes, stream_ptr()),
                 "cbo_ratio_train_step_f64");
    return loss;
}

// in place on X and momentum_buf
void sgld_update(Tensor X, const c10::optional<Tensor>& momentum_buf, const Tensor& dacq_dX, const Tensor& noise, const Tensor& lb,
                 const Tensor& ub, double lr, double momentum, double temperature, double penalty_scale, double penalty_width) {
    c10::cuda::CUDAGuard guard(X.device());
    TORCH_CHECK(X.is_cuda() && X.scalar_type() == torch::kFloat64 && X.is_contiguous(), "X must be a contiguous fp64 CUDA tensor (updated in place)");
    const int64_t d = X.size(-1), n = X.numel() / d;
    Tensor g = dev_f64(dacq_dX, "dacq_dX"), nz = dev_f64(noise, "noise"), l = dev_f64(lb, "lb"), u = dev_f64(ub, "ub");
    double* mb = nullptr;
    if (momentum_buf.has_value() && momentum_buf->defined()) {
        TORCH_CHECK(momentum_buf->is_cuda() && momentum_buf->scalar_type() == torch::kFloat64 && momentum_buf->is_contiguous(),
                    "momentum_buf must be a contiguous fp64 CUDA tensor");
        mb = momentum_buf->data_ptr<double>();
    }
    check_status(cbo_sgld_update_f64(X.data_ptr<double>(), mb, g.data_ptr<double>(), nz.data_ptr<double>(), l.data_ptr<double>(),
                                     u.data_ptr<double>(), n, d, lr, momentum, temperature, penalty_scale, penalty_width, stream_ptr()),
                 "cbo_sgld_update_f64");
}

}  // namespace

TORCH_LIBRARY(cbo, m) {
    m.def(
        "conformal_acq(Tensor X_train, Tensor alpha, Tensor linv, Tensor linv_t, Tensor inv_lengthscale, float mean, float outputscale, "
        "float noise, Tensor Xq, Tensor eps_grid, Tensor z, Tensor eps_mc, Tensor? u_rand, Tensor[] ratio, int kind, float alpha_lvl, "
        "float temp, int grid_res, int max_grid_refinements, bool optimistic, bool randomized, float beta_prime, float best_f, "
        "int chunk) -> Tensor");
    m.def(
        "conformal_acq_raw(Tensor X_train, Tensor alpha, Tensor linv, Tensor linv_t, Tensor inv_lengthscale, float mean, float outputscale, "
        "float noise, Tensor Xq, Tensor eps_grid, Tensor z, Tensor eps_mc, Tensor? u_rand, Tensor[] ratio, int kind, float alpha_lvl, "
        "float temp, int grid_res, int max_grid_refinements, bool optimistic, bool randomized, float beta_prime, float best_f, "
        "int chunk, bool want_grad) -> (Tensor, Tensor)");
    m.def("model_build(Tensor X, Tensor y, Tensor inv_lengthscale, float mean, float outputscale, float noise) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("conf_mask_to_bounds(Tensor target_grid, Tensor conf_pred_mask) -> (Tensor, Tensor)");
    m.def("ratio_forward(Tensor[] weights, Tensor x) -> Tensor");
    m.def(
        "ratio_train_step(Tensor(a!) theta, Tensor(b!) adam_m, Tensor(c!) adam_v, Tensor(d!) target, Tensor x_neg, Tensor x_pos, "
        "Tensor? pos_idx, int n_pos, int step, float lr, float beta1, float beta2, float eps, float weight_decay, float ema_weight) -> Tensor");
    m.def(
        "sgld_update(Tensor(a!) X, Tensor(b!)? momentum_buf, Tensor dacq_dX, Tensor noise, Tensor lb, Tensor ub, float lr, float momentum, "
        "float temperature, float penalty_scale, float penalty_width) -> ()");
}

TORCH_LIBRARY_IMPL(cbo, CUDA, m) {
    m.impl("conformal_acq", &conformal_acq_cuda);
    m.impl("conformal_acq_raw", &conformal_acq_raw);
    m.impl("model_build", &model_build);
    m.impl("conf_mask_to_bounds", &conf_mask_to_bounds);
    m.impl("ratio_forward", &ratio_forward);
    m.impl("ratio_train_step", &ratio_train_step);
    m.impl("sgld_update", &sgld_update);
}

TORCH_LIBRARY_IMPL(cbo, Autograd, m) { m.impl("conformal_acq", &conformal_acq_autograd); }
