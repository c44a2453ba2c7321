"""In-situ cost of each link of the evaluation chain: CUDA-graph replays of growing prefixes
(sqdist | + finalize | + Sinkhorn forward | + Sinkhorn backward | + gradient GEMM and martingale adjoint),
inputs rotated over 5 batches (315 MB > L2).  Differences between consecutive rows are what each link
adds to the critical path with programmatic dependent launch in effect."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import _lib, functional as F
from kccotgan_b200.synthetic import CONFIGS, INPUT_ORDER, make_inputs

lib = _lib.load()
cfg = {k: v for k, v in CONFIGS["cfg2_mazes"].items()}
B, T, J, L, s = cfg["B"], cfg["T"], 8, 100, 1.0 / 15.0
K = cfg["H"] * cfg["W"] * cfg["C"] * T
dev = torch.device("cuda")
p = F._ptr
NS = 5
sets = []
for i in range(NS):
    inp = make_inputs(J=J, kind="uniform", seed=1 + 1000 * i, device=dev, **cfg)
    sets.append([inp[k].contiguous() for k in INPUT_ORDER])
saved = [torch.empty(lib.kccot_mixed_loss_saved_bytes(1, B, L), dtype=torch.uint8, device=dev) for _ in range(NS)]
ws = [torch.empty(lib.kccot_mixed_loss_workspace_bytes(1, B, K, L), dtype=torch.uint8, device=dev) for _ in range(NS)]
out = torch.empty(4, device=dev); gl = torch.ones(1, device=dev)
grads = [[torch.empty_like(t) for t in sets[0]] for _ in range(NS)]
C3 = torch.empty(3, B, B, device=dev)
Cb = torch.empty(3, B, B, device=dev)

def st():
    return F._stream(dev)
def sqdist(i):
    r, f = sets[i][0], sets[i][1]
    _lib.call("kccot_mixed_sqdist_partials", p(r), p(f), 1, B, K, p(ws[i]), ws[i].numel(), 0, st())
def cost(i):
    r, f, hf, mr, hr, mf = sets[i]
    _lib.call("kccot_mixed_cost_fwd", p(r), p(f), 1, B, K, p(hf), p(mr), p(hr), p(mf), T, J, s, p(C3), p(ws[i]), ws[i].numel(), 0, st())
def fwd(i):
    r, f, hf, mr, hr, mf = sets[i]
    _lib.call("kccot_mixed_loss_fwd", p(r), p(f), 1, B, K, p(hf), p(mr), p(hr), p(mf), T, J, s, 1.0, L, p(saved[i]), p(out),
              p(out[1:]), p(ws[i]), ws[i].numel(), 0, st())
def bwd(i, with_real=False):
    r, f, hf, mr, hr, mf = sets[i]
    g = grads[i]
    _lib.call("kccot_mixed_loss_bwd", p(gl), p(r), p(f), 1, B, K, p(hf), p(mr), p(hr), p(mf), T, J, s, 1.0, L, p(saved[i]),
              p(g[0]) if with_real else None, p(g[1]), p(g[2]), p(g[3]), p(g[4]), p(g[5]), p(ws[i]), ws[i].numel(), 0, st())
chains = [("sqdist", lambda i: sqdist(i)), ("+ finalize", lambda i: cost(i)), ("+ sinkhorn fwd (= forward)", lambda i: fwd(i)),
          ("+ backward (full evaluation)", lambda i: (fwd(i), bwd(i))), ("full evaluation incl. d/d real", lambda i: (fwd(i), bwd(i, True)))]
prev = 0.0
for name, fn in chains:
    for i in range(NS): fn(i)
    torch.cuda.synchronize()
    graphs = []
    for i in range(NS):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn(i)
        graphs.append(g)
    for i in range(10): graphs[i % NS].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 500
    e0.record()
    for i in range(reps): graphs[i % NS].replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:36s} {us:7.1f} us  (+{us - prev:5.1f})")
    prev = us
