"""Runs the Sinkhorn kernels alone (for ncu captures and timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import functional as F, _lib
if os.environ.get("KCCOT_LIB_PATH"):          # A/B timing against another build of the library
    _lib.LIB_PATH = os.environ["KCCOT_LIB_PATH"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lib = _lib.load()
torch.manual_seed(0)
C3 = (900 + 4 * torch.randn(n, B, B, device="cuda")).contiguous()
L = int(sys.argv[4]) if len(sys.argv) > 4 else 100
uh = torch.empty(n, L + 1, B, device="cuda"); vh = torch.empty_like(uh)
nits = torch.empty(n, dtype=torch.int32, device="cuda"); cost = torch.empty(n, device="cuda")
ws = torch.empty(max(256, lib.kccot_sinkhorn_workspace_bytes(n, B, L)), dtype=torch.uint8, device="cuda")
g = torch.ones(n, device="cuda"); Cb = torch.empty_like(C3)
st = F._stream(C3.device)
def fwd():
    _lib.call("kccot_sinkhorn_fwd", F._ptr(C3), n, B, 1.0, L, max(L, 100), 1e-2, 0, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(cost), F._ptr(ws), ws.numel(), st)
def bwd():
    _lib.call("kccot_sinkhorn_bwd", F._ptr(C3), n, B, 1.0, L, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(g), F._ptr(Cb), F._ptr(ws), ws.numel(), st)
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} n={n} sinkhorn {name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
print("cost", cost[:3].tolist())
