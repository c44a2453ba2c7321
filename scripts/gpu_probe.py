"""GPU probe (development aid): tcgen05 vs CUDA-core cost kernels, per-kernel timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kccotgan_b200 import functional as F, _lib, gan_utils
from kccotgan_b200.synthetic import make_inputs, INPUT_ORDER, CONFIGS

lib = _lib.load()
print("device", torch.cuda.get_device_name(0), "check", lib.kccot_device_check())


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def cost3(inp, path):
    F.set_path(path)
    r, f = inp["real"], inp["fake"]
    B = r.shape[0]
    R, Fk = r.reshape(B, -1), f.reshape(B, -1)
    K = R.shape[1]
    C3 = torch.empty(3, B, B, device="cuda")
    ws = torch.empty(lib.kccot_mixed_cost_workspace_bytes(1, B, K), dtype=torch.uint8, device="cuda")
    T, J = inp["h_fake"].shape[1:]
    def run():
        _lib.call("kccot_mixed_cost_fwd", F._ptr(R), F._ptr(Fk), 1, B, K, F._ptr(inp["h_fake"]), F._ptr(inp["m_real"]),
                  F._ptr(inp["h_real"]), F._ptr(inp["m_fake"]), T, J, 1 / 15, F._ptr(C3), F._ptr(ws), ws.numel(),
                  F._PATH["flags"], F._stream(r.device))
    run()
    torch.cuda.synchronize()
    return C3, run


for name in sys.argv[1:] or ["cfg1_mmnist", "cfg2_mazes"]:
    c = {k: v for k, v in CONFIGS[name].items() if k != "nprob"}
    inp = {k: v.cuda() for k, v in make_inputs(J=8, kind="uniform", seed=1, **c).items()}
    Cs, run_s = cost3(inp, "simt")
    t_s = timeit(run_s, n=5, warm=1)
    try:
        Ct, run_t = cost3(inp, "tcgen05")
        t_t = timeit(run_t)
        d = (Ct - Cs).abs().max().item()
        print(f"{name}: simt {t_s:.1f} us  tcgen05 {t_t:.1f} us  max|C_tc - C_simt| = {d:.3e}  max|C| = {Cs.abs().max().item():.1f}")
        # reference in fp64 on the GPU (torch, test only)
        R = inp["real"].reshape(c["B"], -1).double(); Fk = inp["fake"].reshape(c["B"], -1).double()
        D = (torch.cdist(R, Fk) ** 2 / 15)
        hm = torch.einsum("itk,jtk->ij", inp["h_fake"][:, :-1].double(), (inp["m_real"][:, 1:] - inp["m_real"][:, :-1]).double()) / 15
        ref = D + hm
        print("   vs fp64: simt", (Cs[0].double() - ref).abs().max().item(), " tc", (Ct[0].double() - ref).abs().max().item())
    except Exception as e:
        print(name, "tcgen05 failed:", e)
    F.set_path("auto")
    leaves = [inp[k].clone().requires_grad_(True) for k in INPUT_ORDER]
    def fb():
        loss = gan_utils.compute_sinkhorn_loss(leaves[0], leaves[1], 1 / 15, 0.8, 100, *leaves[2:], video=True)
        torch.autograd.grad(loss, leaves[1:])
    print(f"   fwd+bwd (auto path): {timeit(fb, n=10):.1f} us")
    # pieces
    B = c["B"]
    C3 = Cs.clone()
    uh = torch.empty(3, 101, B, device="cuda"); vh = torch.empty_like(uh)
    nits = torch.empty(3, dtype=torch.int32, device="cuda"); cost = torch.empty(3, device="cuda")
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    def skf():
        _lib.call("kccot_sinkhorn_fwd", F._ptr(C3), 3, B, 1.0, 100, 100, 1e-2, 0, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(cost), F._ptr(ws), 256, F._stream(C3.device))
    g = torch.ones(3, device="cuda"); Cb = torch.empty_like(C3)
    def skb():
        _lib.call("kccot_sinkhorn_bwd", F._ptr(C3), 3, B, 1.0, 100, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(g), F._ptr(Cb), F._ptr(ws), 256, F._stream(C3.device))
    print(f"   sinkhorn fwd {timeit(skf):.1f} us   bwd {timeit(skb):.1f} us")

    # tcgen05 vs CUDA-core gradients
    outs = {}
    for path in ("simt", "tcgen05"):
        F.set_path(path)
        lv = [inp[k].clone().requires_grad_(True) for k in INPUT_ORDER]
        loss = gan_utils.compute_sinkhorn_loss(lv[0], lv[1], 1 / 15, 0.8, 100, *lv[2:], video=True)
        try:
            outs[path] = torch.autograd.grad(loss, lv[:2])
        except Exception as e:
            print("   grad", path, "failed:", e)
    if len(outs) == 2:
        for nm, a, b in zip(("g_real", "g_fake"), outs["simt"], outs["tcgen05"]):
            print(f"   {nm}: |simt| {a.norm().item():.4e} |tc| {b.norm().item():.4e} rel diff {((a - b).norm() / a.norm()).item():.3e}")
    F.set_path("auto")
