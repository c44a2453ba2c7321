"""torchrun worker of tests/test_sharded_nccl.py: the fused row-sharded mixed loss on N GPUs (NCCL for the one-off
collectives, peer mailboxes inside the persistent Sinkhorn kernels) against the fp64 oracle of the WHOLE problem.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/sharded_nccl_worker.py [B T H W C [reps]]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kccotgan_b200.sharded import ShardedMixedLoss  # noqa: E402
from kccotgan_b200.synthetic import make_inputs  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    B, T, H, W, C = (a + [1024, 4, 16, 16, 1][len(a):])[:5]
    reps = a[5] if len(a) > 5 else 5
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    s = 1.0 / 15.0
    inp = make_inputs(B=B, T=T, H=H, W=W, C=C, J=8, ctx=T // 2, kind="uniform", seed=1)     # same on every rank
    names = ("real", "fake", "h_fake", "m_real", "h_real", "m_fake")
    t = [inp[k].to(dev) for k in names]
    K = T * H * W * C
    sm = ShardedMixedLoss(B, K, T, 8, s, device=dev)
    loss, terms = sm.forward(*t)
    g = sm.backward(1.0)
    torch.cuda.synchronize()
    # timing: forward + backward, device events, max over ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        sm.forward(*t)
        sm.backward(1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # stage breakdown (device events around the backend stages of one more evaluation, this rank)
    be = sm.be
    stages = {}

    def stage(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        stages[name] = a.elapsed_time(b)
        return out
    flat = [t[0].reshape(B, -1).contiguous(), t[1].reshape(B, -1).contiguous()] + [x.contiguous() for x in t[2:]]
    shift = stage("cost_fwd", lambda: be.cost_fwd(*flat))
    stage("allreduce_shift", lambda: sm._allreduce(shift, dist.ReduceOp.MIN))
    part = stage("sinkhorn_fwd", lambda: be.sinkhorn_fwd())
    stage("allreduce_cost", lambda: sm._allreduce(part, dist.ReduceOp.SUM))
    g3 = torch.tensor([2.0, -1.0, -1.0], device=dev)
    Cb = stage("sinkhorn_bwd", lambda: be.sinkhorn_bwd(g3))
    XY = stage("a2a_xy", lambda: sm._columns_of(Cb[0]))
    YY = stage("a2a_yy", lambda: sm._columns_of(Cb[2]))
    stage("cost_bwd", lambda: be.cost_bwd(XY, YY))
    # gather the row-sharded gradients
    full = {}
    for k in ("fake", "h_fake", "h_real"):
        if world > 1:
            parts = [torch.empty_like(g[k]) for _ in range(world)]
            dist.all_gather(parts, g[k].contiguous())
            full[k] = torch.cat(parts, 0)
        else:
            full[k] = g[k]
    if rank == 0:
        if B <= 2048:
            from oracle import closed_form as cf
            npin = [inp[k].numpy() for k in names]
            t0 = time.time()
            ref, gref, det = cf.compute_sinkhorn_loss(npin[0], npin[1], s, 0.8, 100, *npin[2:], video=True, grad=True)
            ref_terms = np.array([det["loss_xy"], det["loss_xx"], det["loss_yy"]])
            scale = np.abs(ref_terms).max()
            terr = np.abs(terms.cpu().numpy().astype(np.float64) - ref_terms).max() / scale
            lerr = abs(float(loss) - ref) / scale
            errs = {}
            for k, n in (("fake", "f_fake"), ("h_fake", "h_fake"), ("h_real", "h_real"), ("m_real", "m_real"), ("m_fake", "m_fake")):
                got = (full[k] if k in full else g[k]).cpu().numpy().astype(np.float64).reshape(gref[n].shape)
                errs[k] = float(np.linalg.norm(got - gref[n]) / np.linalg.norm(gref[n]))
            ok = terr <= 1e-4 and lerr <= 1e-4 and all(e < 1e-4 for e in errs.values())
            print(f"world={world} B={B} K={K}: terms err {terr:.2e} loss err {lerr:.2e} grads {errs} "
                  f"(oracle {time.time() - t0:.1f} s)", flush=True)
        else:
            ok = bool(torch.isfinite(terms).all() and torch.isfinite(full["fake"]).all())
            print(f"world={world} B={B} K={K}: terms {terms.tolist()} (no oracle at this size)", flush=True)
        print(f"fwd+bwd {float(ms):.3f} ms per evaluation (max over ranks)", flush=True)
        print("stages (ms, rank 0):", {k: round(v, 3) for k, v in stages.items()}, flush=True)
        print("SHARDED_OK" if ok else "SHARDED_FAIL", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
