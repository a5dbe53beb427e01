"""Dense stand-ins for LazyTensor / LinearOperator objects (gpytorch 1.9.0): only evaluate / matmul / inv_matmul."""
import torch


def psd_safe_cholesky(M):
    """Plain lower Cholesky; jitter is added only when the factorisation fails."""
    L, info = torch.linalg.cholesky_ex(M)
    if not bool((info != 0).any()):
        return L
    eye = torch.eye(M.shape[-1], dtype=M.dtype, device=M.device)
    for jitter in (1e-8, 1e-7, 1e-6, 1e-5, 1e-4):
        L, info = torch.linalg.cholesky_ex(M + jitter * eye)
        if not bool((info != 0).any()):
            return L
    raise RuntimeError("matrix not positive definite")


class DenseLazy:
    """A matrix (possibly produced on demand by `thunk`) with the LazyTensor calls the reference makes."""

    def __init__(self, tensor=None, thunk=None, shape=None):
        self._t, self._thunk, self._shape = tensor, thunk, shape
        self._chol = None

    def evaluate(self):
        if self._t is None:
            self._t = self._thunk()
        return self._t

    to_dense = evaluate

    @property
    def shape(self):
        return self._t.shape if self._t is not None else torch.Size(self._shape)

    @property
    def batch_shape(self):
        return self.shape[:-2]

    def matmul(self, rhs):
        return self.evaluate() @ rhs

    def cholesky(self):
        if self._chol is None:
            self._chol = psd_safe_cholesky(self.evaluate())
        return self._chol

    def inv_matmul(self, rhs):
        L = self.cholesky()
        return torch.cholesky_solve(rhs.expand(*L.shape[:-2], *rhs.shape[-2:]), L)


class AddedDiagLazy(DenseLazy):
    """K + noise * I with `.lazy_tensors = (K, diag)` (what lik_train_train_covar is; helpers.py:397-398)."""

    def __init__(self, K: DenseLazy, noise):
        self.lazy_tensors = (K, noise)
        n = K.shape[-1]

        def thunk():
            Kd = K.evaluate()
            return Kd + noise[..., None, None] * torch.eye(n, dtype=Kd.dtype, device=Kd.device)

        super().__init__(thunk=thunk, shape=K.shape)


def lazify(t):
    return t if isinstance(t, DenseLazy) else DenseLazy(t)
