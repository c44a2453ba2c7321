"""Development probe: the batched mixed loss (BASELINE config 4 shape) at growing problem counts."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kccotgan_b200 import gan_utils  # noqa: E402

S = 1.0 / 15.0
dev = torch.device("cuda", 0)
B, T, H, W, C = 64, 10, 32, 32, 1
for P in [int(v) for v in (sys.argv[1:] or ["4", "32", "100", "148", "149", "200", "256"])]:
    g = torch.Generator(device=dev).manual_seed(P)
    real = torch.rand((P, B, H, T, W, C), generator=g, device=dev)
    fake = torch.rand((P, B, H, T, W, C), generator=g, device=dev).requires_grad_(True)
    hm = [torch.sigmoid(torch.randn((P, B, T, 8), generator=g, device=dev)).requires_grad_(True) for _ in range(4)]
    ones = torch.ones(P, device=dev)
    print(f"P={P}: forward ...", flush=True)
    t0 = time.time()
    loss = gan_utils.compute_sinkhorn_loss_batched(real, fake, S, *hm)
    torch.cuda.synchronize()
    print(f"   forward done in {time.time() - t0:.3f} s, loss[0]={float(loss[0]):.4f} loss[-1]={float(loss[-1]):.4f}", flush=True)
    t0 = time.time()
    grads = torch.autograd.grad(loss, [fake] + hm, grad_outputs=ones)
    torch.cuda.synchronize()
    print(f"   backward done in {time.time() - t0:.3f} s, |g_fake|={float(grads[0].norm()):.4e}", flush=True)
