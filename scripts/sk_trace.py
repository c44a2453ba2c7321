"""Phase timeline of the small Sinkhorn kernels (clock64 stamps of CTA 0).  Needs a library built with
-DKCCOT_SK_TRACE, e.g. (after `python -m kccotgan_b200.build`):

  F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
  mkdir -p trace_build
  nvcc $F -DKCCOT_SK_TRACE -c kccotgan_b200/csrc/sinkhorn_small.cu -o trace_build/sinkhorn_small.o
  nvcc -shared -o trace_build/libkccot.so $(ls kccotgan_b200/csrc/build/*.o | grep -v sinkhorn_small.o) \
       trace_build/sinkhorn_small.o -lcudart
  KCCOT_LIB_PATH=trace_build/libkccot.so python scripts/sk_trace.py [B] [L]

(with programmatic dependent launch the backward's first phase includes its wait for the forward kernel)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import functional as F, _lib
_lib.LIB_PATH = os.environ["KCCOT_LIB_PATH"]
lib = _lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = 3
torch.manual_seed(0)
C3 = (900 + 4 * torch.randn(n, B, B, device="cuda")).contiguous()
uh = torch.empty(n, L + 1, B, device="cuda"); vh = torch.empty_like(uh)
nits = torch.empty(n, dtype=torch.int32, device="cuda"); cost = torch.empty(n, device="cuda")
ws = torch.empty(max(256, lib.kccot_sinkhorn_workspace_bytes(n, B, L)), dtype=torch.uint8, device="cuda")
g = torch.ones(n, device="cuda"); Cb = torch.empty_like(C3)
st = F._stream(C3.device)
for rep in range(3):
    _lib.call("kccot_sinkhorn_fwd", F._ptr(C3), n, B, 1.0, L, max(L, 100), 1e-2, 0, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(cost), F._ptr(ws), ws.numel(), st)
    _lib.call("kccot_sinkhorn_bwd", F._ptr(C3), n, B, 1.0, L, F._ptr(uh), F._ptr(vh), F._ptr(nits), F._ptr(g), F._ptr(Cb), F._ptr(ws), ws.numel(), st)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 32)()
    raw.kccot_debug_sk_trace(buf)
    t = list(buf)
    names = (["load_slices", "setup (K~)", "iterations", "sharp cost", "history write-out"],
             ["load_slices", "final potentials", "history copy", "seeds", "fast steps", "general steps", "Cbar write"])
    for kern, nm in enumerate(("fwd", "bwd")):
        tt = t[16 * kern: 16 * kern + len(names[kern]) + 1]
        print(nm, "total", tt[-1] - tt[0], "cycles:", ", ".join(f"{a} {tt[j + 1] - tt[j]}" for j, a in enumerate(names[kern])))
