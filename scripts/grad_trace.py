import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import functional as F, _lib
from kccotgan_b200.synthetic import make_inputs, CONFIGS
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
raw.kccot_debug_set_grad_trace.argtypes = [ctypes.c_void_p]
c = {k: v for k, v in CONFIGS["cfg2_mazes"].items() if k != "nprob"}
s = make_inputs(J=8, kind="uniform", seed=1, device="cuda", **c)
B = c["B"]; K = c["T"] * c["H"] * c["W"] * c["C"]; T = c["T"]; J = 8
ws = torch.empty(lib.kccot_mixed_loss_workspace_bytes(1, B, K, 100), dtype=torch.uint8, device="cuda")
Cb = torch.randn(3, B, B, device="cuda") * 1e-3
gf = torch.empty(B, K, device="cuda")
p = F._ptr
def grad():
    _lib.call("kccot_mixed_cost_bwd", p(Cb), p(s["real"]), p(s["fake"]), 1, B, K, p(s["h_fake"]), p(s["m_real"]), p(s["h_real"]), p(s["m_fake"]),
              T, J, 1 / 15, None, p(gf), None, None, None, None, p(ws), ws.numel(), 0, F._stream(ws.device))
grad(); torch.cuda.synchronize()
tr = torch.zeros(7 * 64 * 2, dtype=torch.int64, device="cuda")
raw.kccot_debug_set_grad_trace(ctypes.c_void_p(tr.data_ptr()))
grad(); torch.cuda.synchronize()
raw.kccot_debug_set_grad_trace(None)
t = tr.cpu().view(7, 64, 2)
t0 = t[t > 0].min()
names = ["TMA  (slot free -> issued)", "MMA  (operands ready -> committed)", "CONV (box landed -> converted)", "EPI  (acc ready -> store issued)", "EPI2 (tmem loaded -> wait_read done)", "EPI3 (barrier A passed -> staged)", "EPI4 (fenced -> barrier B passed)"]
for r in range(7):
    print(names[r])
    for i in list(range(0, 6)) + list(range(20, 24)) + list(range(46, 52)):
        a, b = int(t[r, i, 0]), int(t[r, i, 1])
        if a == 0: break
        print(f"   {i:3d}: {a - t0:8d} -> {b - t0:8d}  (+{b - a})")
