"""Does the fp32 Sinkhorn iteration reach a bit-exact fixed point (or short cycle) before L=100?
Reads the potential history kccot_mixed_loss_fwd saves (layout: mixed_abi.cu, C3 | uh | vh | nits | cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import functional as F
from kccotgan_b200.synthetic import CONFIGS, make_inputs

cfg = dict(CONFIGS["cfg2_mazes"]); B, L = cfg["B"], 100
for kind in ("uniform", "video"):
    for seed in (1, 2, 3):
        inp = make_inputs(J=8, kind=kind, seed=seed, device="cuda", **cfg)
        t = [inp[k].requires_grad_(True) for k in ("real", "fake", "h_fake", "m_real", "h_real", "m_fake")]
        loss, terms = F.MixedLossFn.apply(*t, 1 / 15, 1.0, L)
        saved = loss.grad_fn.saved_tensors[-1]
        f = saved.view(torch.int32)
        off = 3 * B * B
        uh = f[off: off + 3 * (L + 1) * B].view(3, L + 1, B)
        vh = f[off + 3 * (L + 1) * B: off + 6 * (L + 1) * B].view(3, L + 1, B)
        for p, nm in enumerate(("xy", "xx", "yy")):
            st = torch.cat([uh[p], vh[p]], 1)
            fix = [k for k in range(1, L) if torch.equal(st[k], st[k + 1])]
            cyc2 = [k for k in range(2, L) if torch.equal(st[k - 1], st[k + 1])]
            print(kind, seed, nm, "first fixed point at", fix[0] if fix else None, "| first 2-cycle at", cyc2[0] if cyc2 else None)
