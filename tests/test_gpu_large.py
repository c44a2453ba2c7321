"""Large-batch path (B > 64: tiled fp16x3 tcgen05 GEMM + streamed / persistent Sinkhorn) against the fp64
oracle, through the reference-shaped Python functions (which call the C ABI).  `pytest -m gpu`.

Same tolerances as test_gpu_parity.py: loss terms 1e-4 of the largest term, gradients 1e-4 rel-L2,
cost matrices 1e-6 of max|C| (2e-6 where noted: GEMM-form in fp32 at |C| ~ 1e3).
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from kccotgan_b200.synthetic import GRAD_NAMES, INPUT_ORDER, make_inputs

pytestmark = pytest.mark.gpu

LOSS_TOL = 1e-4
GRAD_TOL = 1e-4
COST_TOL = 2e-6


@pytest.fixture(scope="module")
def gu():
    from kccotgan_b200 import _lib, gan_utils
    assert _lib.load().kccot_device_check() == 0, _lib.last_error()
    return gan_utils


@pytest.fixture(autouse=True)
def _reset_path():
    from kccotgan_b200 import functional
    functional.set_path("auto")
    yield
    functional.set_path("auto")


def maxrel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g, dtype=torch.float32)


# ---------------------------------------------------------------------------------------------
# cost matrices and their adjoints: self-cost (x is y) and pairs, on and off the aligned shapes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Bx,By,T,D,same", [
    (96, 96, 4, 64, True),        # x == y with 64 < B <= 128 takes the stacked-row tcgen05 kernel
    (128, 128, 4, 64, True),
    (96, 96, 4, 64, False),       # 192 stacked rows: large path
    (160, 160, 5, 200, True),     # large path, symmetric tiles, K = 1000 (not a multiple of 64)
    (160, 200, 5, 200, False),    # large path, rectangular
    (130, 150, 3, 67, False),     # K = 201: odd, unaligned rows (scalar split path)
    (300, 300, 2, 640, True),     # three row tiles, two column tiles of the triangle
])
def test_cost_large(gu, Bx, By, T, D, same):
    from oracle import closed_form as cf
    s = 1.0 / 15.0
    x = _rand((Bx, T, D), 1)
    y = x if same else _rand((By, T, D), 2)
    h = torch.sigmoid(torch.randn((Bx, T, 8), generator=torch.Generator().manual_seed(3)))
    M = torch.sigmoid(torch.randn((By, T, 8), generator=torch.Generator().manual_seed(4)))
    xd = x.cuda().requires_grad_(True)
    yd = xd if same else y.cuda().requires_grad_(True)
    hd, Md = h.cuda().requires_grad_(True), M.cuda().requires_grad_(True)
    C = gu.modified_cost(xd, yd, hd, Md, s)
    ref = cf.modified_cost(x.numpy(), y.numpy(), h.numpy(), M.numpy(), s)
    err = maxrel(C.detach().cpu(), ref)
    print(f"B=({Bx},{By}) K={T * D} same={same}: C max rel err {err:.2e} (max |C| {np.abs(ref).max():.1f})")
    assert err < COST_TOL
    if same:
        assert np.all(np.diag(gu.cost_xy(xd, xd, s).detach().cpu().numpy()) == 0.0)
    # adjoint with a plan-like random Cbar
    Cbar = torch.rand((Bx, By), generator=torch.Generator().manual_seed(5)) ** 8
    Cbar = (Cbar / Cbar.sum()).float()
    leaves = [xd, hd, Md] if same else [xd, yd, hd, Md]
    grads = torch.autograd.grad(C, leaves, grad_outputs=Cbar.cuda())
    Cb = Cbar.numpy().astype(np.float64)
    if same:
        gx_ref = cf.cost_backward(Cb, x.numpy(), x.numpy(), s, same=True)
        refs = [gx_ref]
    else:
        refs = list(cf.cost_backward(Cb, x.numpy(), y.numpy(), s))
    refs += list(cf.martingale_backward(Cb, h.numpy(), M.numpy(), s))
    for name, a, r in zip(("x", "y", "h", "M") if not same else ("x", "h", "M"), grads, refs):
        e = rel_l2(a.cpu().numpy(), r)
        print("   grad", name, f"{e:.2e}")
        assert e < GRAD_TOL, (name, e)


# ---------------------------------------------------------------------------------------------
# the mixed loss above B = 64 (B = 128, 256, 1024 vs closed_form)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T,H,W,C,kind", [
    (128, 6, 16, 16, 2, "uniform"),
    (128, 6, 16, 16, 2, "video"),
    (256, 4, 16, 16, 2, "uniform"),
    (200, 5, 12, 12, 1, "video"),      # B and K off the tile sizes (K = 720)
    (1024, 4, 16, 16, 1, "uniform"),
    (256, 20, 4, 4, 2, "uniform"),     # T * J = 160: config 5's column count in the tiled martingale adjoint (5 groups)
    (256, 30, 4, 4, 1, "video"),       # T * J = 240: all 8 column groups
])
def test_mixed_loss_large(gu, B, T, H, W, C, kind):
    from oracle import closed_form as cf
    s = 1.0 / 15.0
    inp = make_inputs(B=B, T=T, H=H, W=W, C=C, J=8, ctx=T // 2, kind=kind, seed=1)
    leaves = [inp[k].cuda().requires_grad_(True) for k in INPUT_ORDER]
    loss = gu.compute_sinkhorn_loss(leaves[0], leaves[1], s, 0.8, 100, *leaves[2:], video=True)
    grads = torch.autograd.grad(loss, leaves)
    _, terms = gu.sinkhorn_loss_terms(*[t.detach() for t in leaves[:2]], s, *[t.detach() for t in leaves[2:]])
    npin = [inp[k].numpy() for k in INPUT_ORDER]
    rl, rg, det = cf.compute_sinkhorn_loss(npin[0], npin[1], s, 0.8, 100, *npin[2:], video=True, grad=True)
    ref_terms = np.array([det["loss_xy"], det["loss_xx"], det["loss_yy"]])
    scale = np.abs(ref_terms).max()
    t = terms.cpu().numpy().astype(np.float64)
    print(f"B={B} K={T * H * W * C} {kind}: terms err {np.abs(t - ref_terms).max() / scale:.2e}, "
          f"loss err {abs(float(loss) - rl) / scale:.2e}")
    assert np.abs(t - ref_terms).max() <= LOSS_TOL * scale, (t, ref_terms)
    assert abs(float(loss) - rl) <= LOSS_TOL * scale
    for n, a in zip(GRAD_NAMES, grads):
        e = rel_l2(a.cpu().numpy(), rg[n])
        print("   grad", n, f"{e:.2e}")
        assert e < GRAD_TOL, (n, e)


# size-independent properties at a size the oracle cannot reach in seconds: the tensor-core path against the
# CUDA-core direct (x - y)^2 kernels on the device, and exact self-distance zeros
def test_cost_large_vs_direct_form(gu):
    from kccotgan_b200 import functional
    s = 1.0 / 15.0
    B, K = 512, 8192
    x, y = _rand((B, 1, K), 11).cuda(), _rand((B, 1, K), 12).cuda()
    C_tc = gu.cost_xy(x, y, s)
    functional.set_path("simt")
    C_direct = gu.cost_xy(x, y, s)
    functional.set_path("auto")
    err = float((C_tc - C_direct).abs().max() / C_direct.abs().max())
    print("B=512 K=8192: tensor-core vs direct form", err)
    assert err < COST_TOL
    # near-duplicate rows (a generator that copies its input): the distance is a cancellation of large dot products
    y2 = (x + 1e-3 * _rand((B, 1, K), 13).cuda())
    d_tc = gu.cost_xy(x, y2, s).diagonal()
    functional.set_path("simt")
    d_direct = gu.cost_xy(x, y2, s).diagonal()
    functional.set_path("auto")
    err2 = float((d_tc - d_direct).abs().max())
    print("   near-duplicate diagonal: max abs err", err2, "typical C", float(C_direct.abs().max()))
    assert err2 < 1e-6 * float(C_direct.abs().max()) * 50
