import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kccotgan_b200 import _lib, functional as F
lib = _lib.load()
B, eps, L = 24, 0.02, 30
rng = np.random.default_rng(B + L)
C = (900.0 + 4.0 * rng.standard_normal((2, B, B))).astype(np.float32)
C[1] = np.abs(C[1] - 900.0) * 3.0
Ct = torch.from_numpy(C).cuda()
uh = torch.full((2, L + 1, B), 7.0, device="cuda"); vh = torch.full((2, L + 1, B), 7.0, device="cuda")
nits = torch.zeros(2, dtype=torch.int32, device="cuda"); cost = torch.zeros(2, device="cuda")
ws = torch.empty(lib.kccot_sinkhorn_workspace_bytes(2, B, L), dtype=torch.uint8, device="cuda")
_lib.call("kccot_sinkhorn_fwd", F._ptr(Ct), 2, B, eps, L, 100, 1e-2, 0, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(cost), F._ptr(ws), ws.numel(), F._stream(Ct.device))
torch.cuda.synchronize()
print("nits", nits.tolist(), "cost", cost.tolist())
for k in (0, 1, 2, 7, 8, 9, 29, 30):
    print(k, "u", uh[0, k, :4].tolist(), "v", vh[0, k, :4].tolist(), "finite", bool(torch.isfinite(uh[0, k]).all()), bool(torch.isfinite(vh[0, k]).all()))
