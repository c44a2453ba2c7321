import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import functional as F, _lib
lib = _lib.load()
B, T, J, K = 64, 10, 8, 122880
p = F._ptr
Cb = torch.randn(3, B, B, device="cuda") * 1e-3
h = [torch.rand(B, T, J, device="cuda") for _ in range(4)]
g = [torch.empty(B, T, J, device="cuda") for _ in range(4)]
x = torch.rand(B, K, device="cuda"); y = torch.rand(B, K, device="cuda")
ws = torch.empty(lib.kccot_mixed_loss_workspace_bytes(1, B, K, 100), dtype=torch.uint8, device="cuda")
st = F._stream(ws.device)
def mart_only():
    _lib.call("kccot_mixed_cost_bwd", p(Cb), p(x), p(y), 1, B, K, p(h[0]), p(h[1]), p(h[2]), p(h[3]), T, J, 1 / 15, None, None,
              p(g[0]), p(g[1]), p(g[2]), p(g[3]), p(ws), ws.numel(), 0, st)
def single():
    _lib.call("kccot_martingale_bwd", p(Cb), p(h[0]), p(h[1]), 1, B, B, T, J, 1 / 15, p(g[0]), p(g[1]), 0, st)
for nm, fn in (("mixed martingale jobs only", mart_only), ("single pair", single)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(nm, f"{e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call")
