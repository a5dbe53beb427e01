"""SingleTaskGP stand-in: exact GP, ConstantMean + ScaleKernel(MaternKernel(nu=2.5, ARD)) + homoskedastic Gaussian
noise, FIXED hyper-parameters, dense Cholesky.  t > 1 outputs = a batch of independent GPs with
`_aug_batch_shape = (t,)` (botorch BatchedMultiOutputGPyTorchModel layout: train_inputs (t, N, d), train_targets (t, N)).
"""
import math
import types

import torch
from gpytorch.distributions import MultivariateNormal, MultitaskMultivariateNormal
from gpytorch.lazy import AddedDiagLazy, DenseLazy, psd_safe_cholesky

from ..posteriors import GPyTorchPosterior

SQRT5 = math.sqrt(5.0)


def matern52(x1, x2, lengthscale, outputscale):
    """x1 (..., n1, d), x2 (..., n2, d); lengthscale (*hb, 1, d), outputscale (*hb,) broadcast from the right."""
    ls = lengthscale.squeeze(-2)                       # (*hb, d)
    sq = None
    for k in range(x1.shape[-1]):
        lk = ls[..., k][..., None, None]
        dk = (x1[..., :, None, k] - x2[..., None, :, k]) / lk
        sq = dk * dk if sq is None else sq + dk * dk
        del dk
    if torch.is_grad_enabled() and sq.requires_grad:   # under autograd (gradient goldens): no in-place arithmetic
        r = sq.clamp_min(1e-30).sqrt()
        return outputscale[..., None, None] * (torch.exp(-SQRT5 * r) * (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r))
    r = sq.clamp_min_(1e-30).sqrt_()
    out = torch.exp(-SQRT5 * r)
    out *= (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r)
    return outputscale[..., None, None] * out


class _PredictionStrategy:
    def __init__(self, gp):
        self._gp = gp

    @property
    def train_labels(self):
        return self._gp.train_targets

    @property
    def lik_train_train_covar(self):
        gp = self._gp
        return AddedDiagLazy(DenseLazy(gp._K()), gp.likelihood.noise.squeeze(-1))


class SingleTaskGP:
    def __init__(self, train_X, train_Y, *, mean_const, lengthscale, outputscale, noise, _batched=None):
        """train_X (N, d), train_Y (N, t).  Per-output hyper-parameters: mean_const (t,), lengthscale (t, d),
        outputscale (t,), noise (t,) (scalars / (d,) accepted for t = 1)."""
        if _batched is not None:
            self.__dict__.update(_batched)
            return
        t = train_Y.shape[-1]
        f = lambda v, shape: torch.as_tensor(v, dtype=torch.float64).expand(shape).clone()
        self._num_outputs = t
        if t == 1:
            self.train_inputs = (train_X.double(),)
            self.train_targets = train_Y.double().squeeze(-1)
            self._aug_batch_shape = torch.Size([])
            self._mean = f(mean_const, ()).reshape(())
            self._ls = f(lengthscale, (train_X.shape[-1],)).reshape(1, -1)
            self._os = f(outputscale, ()).reshape(())
            noise_t = f(noise, ()).reshape(1)
        else:
            self.train_inputs = (train_X.double().unsqueeze(0).expand(t, -1, -1).contiguous(),)
            self.train_targets = train_Y.double().transpose(-1, -2).contiguous()
            self._aug_batch_shape = torch.Size([t])
            self._mean = f(mean_const, (t,))
            self._ls = f(lengthscale, (t, train_X.shape[-1])).reshape(t, 1, -1)
            self._os = f(outputscale, (t,))
            noise_t = f(noise, (t,)).reshape(t, 1)
        self._input_batch_shape = torch.Size([])
        self.likelihood = types.SimpleNamespace(noise=noise_t)
        self._cache = {}

    # ----- nn.Module-ish surface the reference touches
    def eval(self):
        return self

    def train(self, mode=True):
        self._cache = {}          # "delete pred caches" (helpers.py:187)
        return self

    @property
    def num_outputs(self):
        return self._num_outputs

    @property
    def prediction_strategy(self):
        return _PredictionStrategy(self)

    # ----- exact GP algebra
    def _K(self):
        if "K" not in self._cache:
            X = self.train_inputs[0]
            self._cache["K"] = matern52(X, X, self._ls, self._os)
        return self._cache["K"]

    def _factor(self):
        if "L" not in self._cache:
            K = self._K()
            n = K.shape[-1]
            A = K + self.likelihood.noise.unsqueeze(-1) * torch.eye(n, dtype=K.dtype)
            L = psd_safe_cholesky(A)
            resid = (self.train_targets - self._mean[..., None]).unsqueeze(-1)
            self._cache["L"] = L
            self._cache["alpha"] = torch.cholesky_solve(resid, L).squeeze(-1)
        return self._cache["L"], self._cache["alpha"]

    def _mvn(self, X, observation_noise):
        """X: (*xb, n, d) with xb broadcastable against the model batch shape -> mean (*B, n), lazy cov (*B, n, n)."""
        Xtr = self.train_inputs[0]
        L, alpha = self._factor()
        if X is Xtr:
            Kxs = self._K()
        else:
            Kxs = matern52(X, Xtr, self._ls, self._os)                      # (*B, n, N)
        mean = self._mean[..., None] + (Kxs @ alpha.unsqueeze(-1)).squeeze(-1)
        noise = self.likelihood.noise

        def cov():
            Kxx = self._K() if X is Xtr else matern52(X, X, self._ls, self._os)
            Lb = L.expand(*Kxs.shape[:-2], *L.shape[-2:])
            V = torch.linalg.solve_triangular(Lb, Kxs.transpose(-1, -2), upper=False)
            C = Kxx - V.transpose(-1, -2) @ V
            if observation_noise:
                C = C + noise.unsqueeze(-1) * torch.eye(C.shape[-1], dtype=C.dtype)
            return C

        return MultivariateNormal(mean, DenseLazy(thunk=cov, shape=(*mean.shape, mean.shape[-1])))

    def posterior(self, X, output_indices=None, observation_noise=False, posterior_transform=None, **kwargs):
        if self._num_outputs > 1:
            Xb = X.unsqueeze(-3)                                            # add_output_dim: (*xb, 1, n, d)
            mvn = self._mvn(Xb, observation_noise is not False)            # batch (*xb, t)
            mean, covar = mvn.loc, mvn.lazy_covariance_matrix
            t = self._num_outputs
            mvns = []
            for k in (output_indices or range(t)):
                mvns.append(MultivariateNormal(
                    mean.select(-2, k), DenseLazy(thunk=(lambda k=k: covar.evaluate().select(-3, k)),
                                                  shape=(*mean.shape[:-2], mean.shape[-1], mean.shape[-1]))))
            mvn = MultitaskMultivariateNormal.from_independent_mvns(mvns)
        else:
            mvn = self._mvn(X, observation_noise is not False)
        post = GPyTorchPosterior(mvn)
        if posterior_transform is not None:
            return posterior_transform(post)
        return post

    def condition_on_observations(self, X, Y, **kwargs):
        """Exact GP on the augmented data (what get_fantasy_model's low-rank cache update is equal to, App. B-3)."""
        t = self._num_outputs
        if t > 1:
            inputs = X.unsqueeze(-3).expand(*X.shape[:-2], t, *X.shape[-2:])   # (*xb, t, q, d)
            targets = Y.transpose(-1, -2)                                      # (*xb, t, q)
        else:
            inputs, targets = X, (Y.squeeze(-1) if Y.shape[-1] == 1 else Y)
        Xtr, ytr = self.train_inputs[0], self.train_targets
        bshape = inputs.shape[:-2]
        full_X = torch.cat([Xtr.expand(*bshape, *Xtr.shape[-2:]), inputs], dim=-2)
        full_y = torch.cat([ytr.expand(*bshape, ytr.shape[-1]), targets], dim=-1)
        state = dict(self.__dict__)
        state.update(train_inputs=(full_X,), train_targets=full_y, _cache={},
                     _input_batch_shape=full_y.shape[:(-1 if t == 1 else -2)], _aug_batch_shape=full_y.shape[:-1])
        return SingleTaskGP(None, None, mean_const=None, lengthscale=None, outputscale=None, noise=None, _batched=state)
