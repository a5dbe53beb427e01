"""CPU restatement of the reference's training-side steps around the hot path.  TEST INFRASTRUCTURE:
imported by tests/ only; the product path never calls it.

  ratio_train_step   conformalbo/ratio_estimation.py:110-144 (`optimize_callback` from the minibatch on): BCEWithLogitsLoss
                     (mean) of the classifier d->64->ReLU->64->ReLU->1 over [neg (label 0); pos (label 1)], one
                     torch.optim.Adam step (defaults betas=(0.9, 0.999), eps=1e-8; `weight_decay` is L2 added to the
                     gradient), then `update_target_network` (:101-107).  Written out in numpy (explicit reverse pass and
                     Adam formulas) so that it is independent of torch.optim.
  sgld_update        conformalbo/optim/sgld.py:9-47 with momentum != 0 (SGHMC form), as optimize_acqf_sgld configures it
                     (optim/optimize.py:199-201: momentum=.1, temperature / sqrt(d)).
  cosine_lr          torch.optim.lr_scheduler.CosineAnnealingLR closed form (optim/optimize.py:203-207: T_max=steps//4,
                     eta_min=lr/16), which keeps cycling past T_max.
  bounds_penalty_grad  d/dX of 1e2 * (sigmoid((lb - X)/1e-2).mean() + sigmoid((X - ub)/1e-2).mean())  (optimize.py:228-230)

Pinned against outputs of the reference's own classes run in the build container: tests/golden/ref_training.pt
(oracle/make_golden_training.py); tests/test_oracle_training.py.
"""
import math

import numpy as np


def unpack(theta, d, H=64):
    o = 0
    W1 = theta[o:o + H * d].reshape(H, d); o += H * d
    b1 = theta[o:o + H]; o += H
    W2 = theta[o:o + H * H].reshape(H, H); o += H * H
    b2 = theta[o:o + H]; o += H
    W3 = theta[o:o + H]; o += H
    b3 = theta[o:o + 1]; o += 1
    assert o == theta.size
    return W1, b1, W2, b2, W3, b3


def num_params(d, H=64):
    return H * d + H + H * H + H + H + 1


def classifier_loss_grad(theta, x, labels, d, H=64):
    """Mean BCE-with-logits and its gradient w.r.t. the flat parameter vector (torch's parameter order)."""
    W1, b1, W2, b2, W3, b3 = unpack(theta, d, H)
    n = x.shape[0]
    a1 = x @ W1.T + b1
    h1 = np.maximum(a1, 0.0)
    a2 = h1 @ W2.T + b2
    h2 = np.maximum(a2, 0.0)
    logit = h2 @ W3 + b3[0]
    # max(l, 0) - l y + log1p(exp(-|l|))
    loss = np.mean(np.maximum(logit, 0.0) - logit * labels + np.log1p(np.exp(-np.abs(logit))))
    dl = (1.0 / (1.0 + np.exp(-logit)) - labels) / n
    gW3 = h2.T @ dl
    gb3 = np.array([dl.sum()])
    dz2 = np.outer(dl, W3) * (a2 > 0)
    gW2 = dz2.T @ h1
    gb2 = dz2.sum(0)
    dz1 = (dz2 @ W2) * (a1 > 0)
    gW1 = dz1.T @ x
    gb1 = dz1.sum(0)
    return loss, np.concatenate([gW1.ravel(), gb1, gW2.ravel(), gb2, gW3, gb3])


def ratio_train_step(theta, m, v, target, step, neg, pos, d, lr, weight_decay, ema_weight,
                     beta1=0.9, beta2=0.999, eps=1e-8, H=64):
    """One `optimize_callback` training step; `step` is 1-based.  Returns (theta, m, v, target, loss), all new arrays."""
    x = np.concatenate([neg, pos], 0)
    labels = np.concatenate([np.zeros(len(neg)), np.ones(len(pos))])
    loss, g = classifier_loss_grad(theta, x, labels, d, H)
    g = g + weight_decay * theta
    m = m + (1.0 - beta1) * (g - m)
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    theta = theta - (lr / bc1) * m / denom
    target = target * (1.0 - ema_weight) + ema_weight * theta
    return theta, m, v, target, loss


def ratio_forward(target, x, d, H=64):
    W1, b1, W2, b2, W3, b3 = unpack(target, d, H)
    h = np.maximum(x @ W1.T + b1, 0.0)
    h = np.maximum(h @ W2.T + b2, 0.0)
    p = 1.0 / (1.0 + np.exp(-(h @ W3 + b3[0])))
    return np.minimum(p, 1 - 1e-6) / np.maximum(1 - p, 1e-6)


def sgld_update(x, buf, grad, noise, lr, momentum, temperature):
    buf = buf * (1.0 - momentum) - lr * grad + math.sqrt(2.0 * lr * momentum * temperature) * noise
    return x + buf, buf


def cosine_lr(step, lr0, T_max, eta_min):
    """Learning rate in force at 0-based `step` (closed form of CosineAnnealingLR; periodic with period 2 T_max)."""
    return eta_min + (lr0 - eta_min) * (1.0 + math.cos(math.pi * step / T_max)) / 2.0


def bounds_penalty_grad(X, lb, ub, scale=1e2, width=1e-2):
    n = X.size
    sl = 1.0 / (1.0 + np.exp(-(lb - X) / width))
    su = 1.0 / (1.0 + np.exp(-(X - ub) / width))
    return scale * (-sl * (1 - sl) + su * (1 - su)) / (width * n)
