"""Runs the tcgen05 cost kernels alone on a BASELINE config (for ncu captures and timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import functional as F, _lib
from kccotgan_b200.synthetic import make_inputs, CONFIGS
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_mazes"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lib = _lib.load()
c = {k: v for k, v in CONFIGS[name].items() if k != "nprob"}
sets = []
for i in range(5):
    inp = make_inputs(J=8, kind="uniform", seed=1 + i, device="cuda", **c)
    sets.append(inp)
B = c["B"]; K = c["T"] * c["H"] * c["W"] * c["C"]; T = c["T"]; J = 8
ws = torch.empty(lib.kccot_mixed_loss_workspace_bytes(1, B, K, 100), dtype=torch.uint8, device="cuda")
Cb = torch.randn(3, B, B, device="cuda") * 1e-3
gf = torch.empty(B, K, device="cuda")
st = F._stream(ws.device); p = F._ptr
def part(i):
    s = sets[i % 5]
    _lib.call("kccot_mixed_sqdist_partials", p(s["real"]), p(s["fake"]), 1, B, K, p(ws), ws.numel(), 0, st)
def grad(i):
    s = sets[i % 5]
    _lib.call("kccot_mixed_cost_bwd", p(Cb), p(s["real"]), p(s["fake"]), 1, B, K, p(s["h_fake"]), p(s["m_real"]), p(s["h_real"]), p(s["m_fake"]),
              T, J, 1 / 15, None, p(gf), None, None, None, None, p(ws), ws.numel(), 0, st)
for nm, fn in (("sqdist partials", part), ("grad (fake only)", grad)):
    for i in range(2): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i + 2)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} {nm}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per call (back-to-back)")
