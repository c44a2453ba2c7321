"""Row-sharded Sinkhorn over NCCL (kccotgan_b200.sharded.ShardedSinkhorn) on N GPUs against the
unsharded streamed solve of the same cost on rank 0.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/sharded_nccl_check.py [B] [L]
Also times one forward+backward (device events, max over ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from kccotgan_b200 import _lib, functional as F
from kccotgan_b200.sharded import CudaShardBackend, ShardedSinkhorn, row_range

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(11)
C = (900.0 + 4.0 * torch.randn((B, B), generator=g)).to(dev)          # same matrix on every rank
r0, r1 = row_range(B, rank, world)
sk = ShardedSinkhorn(CudaShardBackend(C[r0:r1].contiguous(), B, 1.0))
cost = sk.forward(L=L)
Cbar_rows = sk.backward(g=2.0)
torch.cuda.synchronize()
# timing
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(3):
    sk.forward(L=L); sk.backward(g=2.0)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# reference: the unsharded streamed path on rank 0
gathered = [torch.empty((row_range(B, r, world)[1] - row_range(B, r, world)[0], B), device=dev) for r in range(world)]
dist.all_gather(gathered, Cbar_rows.contiguous())
if rank == 0:
    lib = _lib.load()
    uh = torch.empty(1, L + 1, B, device=dev); vh = torch.empty_like(uh)
    nits = torch.empty(1, dtype=torch.int32, device=dev); c1 = torch.empty(1, device=dev)
    ws = torch.empty(lib.kccot_sinkhorn_workspace_bytes(1, B, L), dtype=torch.uint8, device=dev)
    gc = torch.full((1,), 2.0, device=dev); Cb = torch.empty(1, B, B, device=dev)
    st = F._stream(dev); p = F._ptr
    _lib.call("kccot_sinkhorn_fwd", p(C), 1, B, 1.0, L, 100, 1e-2, 0, p(uh), p(vh), p(nits), p(c1), p(ws), ws.numel(), st)
    _lib.call("kccot_sinkhorn_bwd", p(C), 1, B, 1.0, L, p(uh), p(vh), p(nits), p(gc), p(Cb), p(ws), ws.numel(), st)
    torch.cuda.synchronize()
    full = torch.cat(gathered, 0)
    rel = float((full - Cb[0]).norm() / Cb[0].norm())
    dc = abs(float(cost) - float(c1[0])) / abs(float(c1[0]))
    print(f"sharded x{world}: B={B} L={L} cost {float(cost):.6f} vs unsharded {float(c1[0]):.6f} (rel {dc:.2e}); "
          f"Cbar rel-L2 {rel:.2e}; fwd+bwd {float(ms):.2f} ms (max over ranks)")
    assert dc < 1e-4 and rel < 1e-4          # the 1e-4 fp32 parity bar of the path (different summation orders)
dist.destroy_process_group()
