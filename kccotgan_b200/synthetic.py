"""Seeded synthetic inputs of the shapes the loss path sees.

All values are drawn on the CPU generator in fp32 (so every box reproduces the same bits) and
then moved to `device`.  Layouts follow the reference's training step: videos `[B,H,T,W,C]`
(kernel_train.py:301-306), discriminator outputs `[B,T,J]` in the sigmoid range (gan.py:418).
"""
import torch

# (B, T, ctx, H, W, C) of the BASELINE.json configs
CONFIGS = {
    "cfg1_mmnist": dict(B=32, T=20, ctx=10, H=64, W=64, C=1),
    "cfg2_mazes": dict(B=64, T=10, ctx=3, H=64, W=64, C=3),
    "cfg3_bair": dict(B=64, T=12, ctx=2, H=64, W=64, C=3),
    "cfg4_batched": dict(B=64, T=10, ctx=3, H=32, W=32, C=1, nprob=256),
    "cfg5_large": dict(B=8192, T=20, ctx=10, H=64, W=64, C=1),
}


def make_inputs(B, T, H, W, C, J=8, ctx=None, kind="uniform", seed=1, device="cpu",
                dtype=torch.float32):
    """Returns dict(real, fake, h_fake, m_real, h_real, m_fake).

    kind="uniform": real, fake ~ U[0,1) i.i.d. (numerically the hard case: large, nearly equal costs).
    kind="video":   real = Bernoulli(0.1)*U[0,1); fake shares the first `ctx` frames with real
                    (kernel_train.py:268) and is clip(real + 0.1*N(0,1), 0, 1) on the rest.
    """
    f32 = torch.float32   # draws are always fp32, whatever the default dtype
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    shape = (B, H, T, W, C)
    if kind == "uniform":
        real = torch.rand(shape, generator=g, dtype=f32)
        fake = torch.rand(shape, generator=g, dtype=f32)
    elif kind == "video":
        ctx = T // 2 if ctx is None else ctx
        real = (torch.rand(shape, generator=g, dtype=f32) < 0.1).to(f32) * torch.rand(shape, generator=g, dtype=f32)
        fake = (real + 0.1 * torch.randn(shape, generator=g, dtype=f32)).clamp_(0.0, 1.0)
        fake[:, :, :ctx] = real[:, :, :ctx]
    else:
        raise ValueError(kind)
    out = {"real": real, "fake": fake}
    for name in ("h_fake", "m_real", "h_real", "m_fake"):
        out[name] = torch.sigmoid(torch.randn((B, T, J), generator=g, dtype=f32))
    return {k: v.to(device=device, dtype=dtype) for k, v in out.items()}


INPUT_ORDER = ("real", "fake", "h_fake", "m_real", "h_real", "m_fake")
GRAD_NAMES = ("f_real", "f_fake", "h_fake", "m_real", "h_real", "m_fake")
