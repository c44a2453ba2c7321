"""Development probe: CUDA-graph replay of the batched mixed loss; argv: P mode(fwd|both)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kccotgan_b200 import gan_utils  # noqa: E402

S = 1.0 / 15.0
dev = torch.device("cuda", 0)
B, T, H, W, C = 64, 10, 32, 32, 1
P = int(sys.argv[1])
mode = sys.argv[2] if len(sys.argv) > 2 else "both"
g = torch.Generator(device=dev).manual_seed(P)
real = torch.rand((P, B, H, T, W, C), generator=g, device=dev)
fake = torch.rand((P, B, H, T, W, C), generator=g, device=dev).requires_grad_(True)
hm = [torch.sigmoid(torch.randn((P, B, T, 8), generator=g, device=dev)).requires_grad_(True) for _ in range(4)]
ones = torch.ones(P, device=dev)


def step():
    loss = gan_utils.compute_sinkhorn_loss_batched(real, fake, S, *hm)
    if mode == "fwd":
        return loss, None
    return loss, torch.autograd.grad(loss, [fake] + hm, grad_outputs=ones)


side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    for _ in range(2):
        step()
torch.cuda.current_stream(dev).wait_stream(side)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    out = step()
print(f"P={P} {mode}: captured", flush=True)
for i in range(3):
    gr.replay()
    torch.cuda.synchronize()
    print(f"   replay {i} ok, loss[0]={float(out[0][0].detach()):.4f}", flush=True)
