import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200 import _lib, functional as F
lib = ctypes.CDLL(_lib.LIB_PATH)
fn = lib.kccot_debug_tma_stream
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
K = 122880
sink = torch.zeros(1, device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for rows in (128, 64):
    xs = [torch.rand(rows, K, device="cuda") for _ in range(4)]
    for stages in (4, 8, 12):
        for pf in (0, 16, 32):
            if stages * rows * 128 > 200 * 1024: continue
            ts = []
            for i in range(6):
                x = xs[i % 4]
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(x.data_ptr(), rows, K, stages, pf, sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
                e1.record(); torch.cuda.synchronize()
                assert rc == 0, _lib.last_error()
                ts.append(e0.elapsed_time(e1) * 1e3)
            t = sorted(ts)[1]
            print(f"rows={rows} stages={stages} prefetch={pf}: {t:.1f} us  {rows * K * 4 / t / 1e3:.0f} GB/s")
