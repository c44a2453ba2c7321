"""Development probe: where the host time of the eager reference-signature call goes (cProfile, config 2)."""
import cProfile
import os
import pstats
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kccotgan_b200 import gan_utils  # noqa: E402
from kccotgan_b200.synthetic import CONFIGS, INPUT_ORDER, make_inputs  # noqa: E402
dev = torch.device("cuda", 0)
cfg = {k: v for k, v in CONFIGS["cfg2_mazes"].items() if k != "nprob"}
sets = []
for i in range(5):
    inp = make_inputs(J=8, kind="uniform", seed=1 + 1000 * i, device=dev, **cfg)
    sets.append([inp[k].requires_grad_(k != "real") for k in INPUT_ORDER])
def step(lv):
    loss = gan_utils.compute_sinkhorn_loss(lv[0], lv[1], 1 / 15, 0.8, 100, *lv[2:], video=True)
    return loss, torch.autograd.grad(loss, lv[1:])
for i in range(20):
    step(sets[i % 5])
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for i in range(n):
    step(sets[i % 5])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host submit {1e6 * (t1 - t0) / n:.1f} us per eval, until done {1e6 * (t2 - t0) / n:.1f} us per eval")
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    step(sets[i % 5])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
