"""The reference's formulation of the loss path, op for op, on torch CPU tensors.

TEST INFRASTRUCTURE ONLY.  This is what `bench.py` times as the CPU baseline (`cpu_baseline.kind =
"port"`, and the `--impl reference` arm) on the GPU box, where `/root/reference` and TensorFlow
do not exist: the same eager op sequence the reference issues — the `[B,B,T,D]` broadcast
difference (gan_utils.py:14-17), the `[B,B,T-1,J]` martingale product (:34-38), 2·L row
logsumexps with the potentials re-added (:151-156), the sharp cost (:162-164) — differentiated
by autograd through the unrolled loop exactly as TF's GradientTape does.  dtype follows the
inputs (fp32 for timing, fp64 when used as a cross-check of `closed_form.py`).
"""
import math

import torch


def pairwise_sq_cost(x, y, s):
    # gan_utils.py:14-17
    diff = x.unsqueeze(1) - y.unsqueeze(0)
    return (diff ** 2).sum(-1).sum(-1) * s


def causal_cost(x, y, h, M, s):
    # gan_utils.py:34-43 — h by row sample, Delta M by column sample
    dM = M[:, 1:, :] - M[:, :-1, :]
    ht = h[:, :-1, :]
    hm = (ht[:, None, :, :] * dM[None, :, :, :]).sum(-1).sum(-1) * s
    return pairwise_sq_cost(x, y, s) + hm


def sinkhorn_sharp_cost(C, eps=1.0, L=100, Lmin=100, thresh=1e-2):
    # gan_utils.py:138-165
    n = C.shape[0]
    logmu = torch.full((n, 1), math.log(1.0 / n), dtype=C.dtype)
    u = torch.zeros(n, 1, dtype=C.dtype)
    v = torch.zeros(n, 1, dtype=C.dtype)
    done = 0
    for _ in range(L):
        u_prev = u
        K = (-C + u + v.t()) / eps
        u = eps * (logmu - torch.logsumexp(K, dim=1, keepdim=True)) + u
        K = (-C + u + v.t()) / eps
        v = eps * (logmu - torch.logsumexp(K.t(), dim=1, keepdim=True)) + v
        done += 1
        if thresh > float((u - u_prev).abs().sum()) and done >= Lmin:
            break
    plan = torch.exp((-C + u + v.t()) / eps)
    return (plan * C).sum()


def mixed_loss(f_real, f_fake, s, h_fake, m_real, h_real, m_fake, video=True):
    # gan_utils.py:216-225; eps/L are fixed at 1.0/100 by the positional-argument slip (:221-223)
    if video:
        f_real = f_real.permute(0, 2, 1, 3, 4).reshape(f_real.shape[0], f_real.shape[2], -1)
        f_fake = f_fake.permute(0, 2, 1, 3, 4).reshape(f_fake.shape[0], f_fake.shape[2], -1)
    xy = sinkhorn_sharp_cost(causal_cost(f_real, f_fake, h_fake, m_real, s))
    xx = sinkhorn_sharp_cost(causal_cost(f_real, f_real, h_real, m_real, s))
    yy = sinkhorn_sharp_cost(causal_cost(f_fake, f_fake, h_fake, m_fake, s))
    return 2.0 * xy - xx - yy


def mixed_loss_fwd_bwd(f_real, f_fake, s, h_fake, m_real, h_real, m_fake, video=True):
    """One metric 'eval': loss value + gradients w.r.t. f_fake, h_fake, m_real, h_real, m_fake."""
    leaves = [t.detach().clone().requires_grad_(True) for t in (f_fake, h_fake, m_real, h_real, m_fake)]
    ff, hf, mr, hr, mf = leaves
    loss = mixed_loss(f_real, ff, s, hf, mr, hr, mf, video)
    grads = torch.autograd.grad(loss, leaves)
    return loss.detach(), dict(zip(("f_fake", "h_fake", "m_real", "h_real", "m_fake"), grads))
