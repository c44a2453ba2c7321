"""Times the smoothing kernels (KernelSmoothing.temporal_convolution / gaussian_convolution3D, forward and
backward) at BASELINE config 3's video shape; prints us and GB/s against the algorithmic bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kccotgan_b200.data_utils import KernelSmoothing
from kccotgan_b200.synthetic import CONFIGS

c = CONFIGS["cfg3_bair"]
B, T, H, W, C = c["B"], c["T"], c["H"], c["W"], c["C"]
nbytes = B * T * H * W * C * 4
ks = KernelSmoothing(temporal_kernel_size=6, spatial_kernel_size=6)
xs = [torch.rand(B, H, T, W, C, device="cuda").requires_grad_(True) for _ in range(6)]     # rotate: 6 x 151 MB > L2
def timeit(fn, reps=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for name, op in (("temporal_convolution", ks.temporal_convolution), ("gaussian_convolution3D", ks.gaussian_convolution3D)):
    tf = timeit(lambda i: op(xs[i % 6], 5.0))
    outs = [op(x, 5.0) for x in xs]
    gs = [torch.rand_like(o) for o in outs]
    tb = timeit(lambda i: torch.autograd.grad(outs[i % 6], xs[i % 6], gs[i % 6], retain_graph=True))
    # the same C-ABI calls captured in CUDA graphs (no Python / launch gaps; input L2-resident)
    from kccotgan_b200 import _lib, functional as F
    lib = _lib.load()
    mode = 1 if name == "temporal_convolution" else 3
    dev = xs[0].device
    taps, rad = F._taps(tuple(ks._weights(3, 5.0)))
    x0 = xs[0].detach(); g0 = gs[0]
    o0 = torch.empty_like(x0); gx0 = torch.empty_like(x0); mx = torch.empty(1, device=dev)
    wsb = torch.empty(lib.kccot_smooth_workspace_bytes(mode, B, H, T, W, C), dtype=torch.uint8, device=dev)
    p = F._ptr
    def fwd_abi():
        _lib.call("kccot_smooth_fwd", mode, p(x0), B, H, T, W, C, taps, rad, taps, rad, p(o0), p(mx), p(wsb), wsb.numel(),
                  F._stream(dev))
    def bwd_abi():
        _lib.call("kccot_smooth_bwd", mode, p(g0), p(o0), p(mx), B, H, T, W, C, taps, rad, taps, rad, p(gx0), p(wsb),
                  wsb.numel(), F._stream(dev))
    fwd_abi(); bwd_abi(); torch.cuda.synchronize()
    gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gf):
        fwd_abi()
    with torch.cuda.graph(gb):
        bwd_abi()
    tgf = timeit(lambda i: gf.replay()); tgb = timeit(lambda i: gb.replay())
    print(f"{name}: graph replay fwd {tgf:.1f} us, bwd {tgb:.1f} us (input L2-resident)")
    print(f"{name}: tensor {nbytes / 1e6:.1f} MB; fwd {tf:.1f} us ({2 * nbytes / tf / 1e3:.0f} GB/s on read+write), "
          f"bwd {tb:.1f} us ({2 * nbytes / tb / 1e3:.0f} GB/s on read+write)")
