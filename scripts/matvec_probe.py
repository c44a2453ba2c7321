"""Cycles per Sinkhorn scaling iteration (B=64) for 4, 2 and 1 lanes per row (kccot_debug_matvec_probe)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import _lib
_lib.load()
raw = ctypes.CDLL(_lib.LIB_PATH)
C = (900 + 4 * torch.randn(64, 64, device="cuda")).contiguous()
cyc = torch.zeros(1, dtype=torch.int64, device="cuda"); ab = torch.zeros(3 * 128, device="cuda")
res = {}
for lpr in (4, 2, 1):
    for iters in (100, 300):
        for _ in range(2):
            rc = raw.kccot_debug_matvec_probe(lpr, iters, ctypes.c_void_p(C.data_ptr()), ctypes.c_void_p(cyc.data_ptr()),
                                              ctypes.c_void_p(ab.data_ptr()), None)
            torch.cuda.synchronize()
        res[(lpr, iters)] = int(cyc[0])
    per = (res[(lpr, 300)] - res[(lpr, 100)]) / 200
    print(f"lanes/row {lpr}: {per:.0f} cycles per iteration ({per / 1.965:.0f} ns at 1965 MHz); a[0..2] = {ab[:3].tolist()}")
