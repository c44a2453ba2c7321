#!/usr/bin/env python
"""bench.py — causal-OT loss forward+backward evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one eval = `compute_sinkhorn_loss` value + gradients w.r.t. f_fake, h_fake, m_real,
h_real, m_fake on synthetic inputs of BASELINE config 2 (GQN Mazes: B=64, 3+7 frames
of 64x64x3, J=8).  With N GPUs every rank evaluates its own independent batch (problem-parallel,
no data-path collective: SURVEY.md §8e) -> weak scaling, value = N*K / max-over-ranks time.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, timed with CUDA events, inputs
rotated over more data than L2 holds.  `e2e`: the same call with HOST (pinned) inputs, H2D copies
and the D2H read of the loss inside the timed region.  `roofline`: the dominant HBM-bound kernel.
`cpu_baseline`: the reference formulation (oracle/port_torch.py) on the box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S = 1.0 / 15.0
METRIC = "causal_ot_loss_fwd_bwd_evals_per_sec"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_mazes")
    ap.add_argument("--kind", default="uniform", choices=["uniform", "video"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time eager Python calls instead of CUDA-graph replays")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def workload_config(name, kind):
    from kccotgan_b200.synthetic import CONFIGS
    c = dict(CONFIGS[name])
    c.pop("nprob", None)
    K = c["T"] * c["H"] * c["W"] * c["C"]
    return c, K, {"workload": f"{name}: B={c['B']} T={c['T']} ({c['ctx']}+{c['T'] - c['ctx']}) frames "
                              f"{c['H']}x{c['W']}x{c['C']} J=8 s=1/15 eps=1.0 L=100 ({kind} inputs)",
                  "B": c["B"], "K": K, "T": c["T"], "J": 8, "inputs": kind}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_reference_timing(cfg, kind, steps, warmup, budget_s):
    """Times the reference formulation (oracle/port_torch.py: [B,B,T,D] broadcast cost, eager
    logsumexp loop, autograd through the unrolled iterations) in fp32 on all host cores."""
    import torch
    from kccotgan_b200.synthetic import make_inputs
    from oracle import port_torch as pt
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp = make_inputs(J=8, kind=kind, seed=1, **cfg)
    args = (inp["real"], inp["fake"], S, inp["h_fake"], inp["m_real"], inp["h_real"], inp["m_fake"])
    t_start = time.perf_counter()
    for _ in range(max(0, min(warmup, 1))):
        pt.mixed_loss_fwd_bwd(*args)
    done, elapsed = 0, 0.0
    while done < max(1, steps):
        t0 = time.perf_counter()
        pt.mixed_loss_fwd_bwd(*args)
        elapsed += time.perf_counter() - t0
        done += 1
        if time.perf_counter() - t_start > budget_s:
            break
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"value": done / elapsed, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} full evals of the workload (1 warm-up), fp32, torch-CPU {cores} threads on '{model}'; "
                      "reference formulation restated in oracle/port_torch.py (TensorFlow and /root/reference are "
                      "absent on the GPU box)",
            "s_per_eval": elapsed / done}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, K, config = workload_config(args.workload, args.kind)
    r = cpu_reference_timing(cfg, args.kind, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["s_per_eval"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config, "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for t, l in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        for l in rows:
            parts = [p.strip() for p in l.split(",")]
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from kccotgan_b200 import _lib, functional as F, gan_utils
    from kccotgan_b200.synthetic import INPUT_ORDER, make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kccotgan_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.kccot_device_check())
    cfg, K, config = workload_config(args.workload, args.kind)
    B, T = cfg["B"], cfg["T"]

    # ---- inputs: NSETS distinct batches so that consecutive steps never find their videos in L2
    in_bytes = 2 * B * K * 4
    nsets = max(3, int(300e6 // in_bytes) + 1)
    sets = []
    for i in range(nsets):
        inp = make_inputs(J=8, kind=args.kind, seed=1 + rank + 1000 * i, device=dev, **cfg)
        sets.append([inp[k].requires_grad_(k != "real") for k in INPUT_ORDER])
    config["l2_policy"] = f"inputs rotated over {nsets} distinct batches ({nsets * in_bytes / 1e6:.0f} MB > 126 MB L2)"
    config["parallelism"] = f"problem-parallel x{world} (independent batch per GPU, no collective)"

    def step(leaves):
        loss = gan_utils.compute_sinkhorn_loss(leaves[0], leaves[1], S, 0.8, 100, *leaves[2:], video=True)
        grads = torch.autograd.grad(loss, leaves[1:])
        return loss, grads

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(sets[i % nsets])
    barrier()
    # one captured graph per input set (static-buffer contract of kccotgan_b200.graphed)
    from kccotgan_b200.graphed import GraphedSinkhornLoss
    graphs = None
    if not args.eager:
        graphs = [GraphedSinkhornLoss(*[t.detach() for t in sets[i]], S, adopt=True) for i in range(nsets)]
        for i in range(max(3, args.warmup)):
            graphs[i % nsets].step()
        barrier()
        # the replayed graph must reproduce the eager result bit for bit
        l_e, g_e = step(sets[0])
        graphs[0].step()
        torch.cuda.synchronize()
        assert float(l_e) == float(graphs[0].loss), (float(l_e), float(graphs[0].loss))
        assert torch.equal(g_e[0], graphs[0].grads["fake"])
        config["launch"] = "CUDA-graph replay of the fused forward+backward chain (kccotgan_b200.graphed)"
    else:
        config["launch"] = "eager Python calls (gan_utils.compute_sinkhorn_loss + torch.autograd.grad)"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.kccot_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        if graphs is not None:
            graphs[i % nsets].step()
        else:
            step(sets[i % nsets])
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = lib.kccot_launch_count() - launches0
    if graphs is not None:
        launches = args.steps * graphs[0].kernels_per_replay      # replays do not pass through the C-ABI counter
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = world * args.steps / (ms * 1e-3)
    eager_evals = None
    if graphs is not None:                          # also report the eager Python path
        ne = max(10, min(args.steps, 200))
        barrier()
        e0.record()
        for i in range(ne):
            step(sets[i % nsets])
        e1.record()
        barrier()
        eager_evals = world * ne / (e0.elapsed_time(e1) * 1e-3)

    # ---- e2e: host (pinned) inputs through the public API, H2D + D2H inside the timed region
    host = [[t.detach().cpu().pin_memory() for t in sets[i]] for i in range(min(nsets, 3))]
    h2d = sum(t.numel() * 4 for t in host[0])

    # Double-buffered: the H2D copy of step i+1 runs on a copy stream while step i computes; the loss of
    # step i is copied to pinned host memory asynchronously and read one step later (what a training loop
    # that prefetches its next batch and logs its loss does).  Every step still copies its own inputs.
    copy_s = torch.cuda.Stream(dev)
    main_s = torch.cuda.current_stream(dev)
    dbuf = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host[0]] for _ in range(2)]
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_loss = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        loss_pin = torch.empty(n, dtype=torch.float32).pin_memory()
        out = []
        for i in range(n):
            b = i % 2
            with torch.cuda.stream(copy_s):
                if i >= 2:
                    copy_s.wait_event(ev_free[b])          # the compute that read this buffer two steps ago is done
                for d, h in zip(dbuf[b], host[i % len(host)]):
                    d.copy_(h, non_blocking=True)
                ev_copied[b].record(copy_s)
            main_s.wait_event(ev_copied[b])
            leaves = [d.detach().requires_grad_(j != 0) for j, d in enumerate(dbuf[b])]
            loss, grads = step(leaves)
            loss_pin[i].copy_(loss.detach(), non_blocking=True)     # D2H read of the loss
            ev_loss[b].record(main_s)
            ev_free[b].record(main_s)
            if i >= 1:
                ev_loss[(i - 1) % 2].synchronize()
                out.append(float(loss_pin[i - 1]))
        ev_loss[(n - 1) % 2].synchronize()
        out.append(float(loss_pin[n - 1]))
        return out

    e2e_run(4)
    barrier()
    n_e2e = max(5, min(args.steps, 40))
    e0.record()
    e2e_losses = e2e_run(n_e2e)
    e1.record()
    barrier()
    assert all(math.isfinite(v) for v in e2e_losses)
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms)
    e2e = {"value": world * n_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "steps": n_e2e, "pipeline": "H2D of step i+1 on a copy stream overlaps the compute of step i (2 device "
                                        "buffers); loss read back one step late"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-stage device times + roofline of the dominant HBM kernel (rank 0, after the timed region)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    stages = stage_times(lib, F, torch, sets, B, K, T, dev)
    # algorithmic bytes: forward distances read X,Y once (8BK); adjoint reads X,Y and writes g_fake (12BK)
    alg = {"sqdist_tc_kernel": 8.0 * B * K, "grad_tc_kernel": 12.0 * B * K}
    traffic = {}
    try:        # DRAM bytes per launch from the committed `ncu --set full` capture of the same kernels
        with open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json")) as f:
            traffic = json.load(f)
    except OSError:
        pass
    dom = max(alg, key=lambda k: stages.get(k, 0.0))
    dur_us = stages[dom]
    achieved = alg[dom] / (dur_us * 1e-6) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic.get(dom.split("(")[0]), "peak_source": peak_src,
                "other_hbm_kernels": {k: {"achieved_GBps": alg[k] / (stages[k] * 1e-6) / 1e9,
                                          "frac": alg[k] / (stages[k] * 1e-6) / 1e9 / hbm_peak} for k in alg if k != dom},
                "algorithmic_bytes_per_launch": alg[dom], "avg_launch_us": dur_us,
                "stage_us": stages,
                "whole_eval": {"algorithmic_bytes": 20.0 * B * K + 40.0 * B * T * 8,
                               "frac_of_hbm_roofline": (20.0 * B * K + 40.0 * B * T * 8) / (ms * 1e-3 / args.steps) / 1e9
                               / hbm_peak}}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_timing(cfg, args.kind, steps=3, warmup=1, budget_s=args.cpu_budget_s)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (cost GEMMs 3xTF32 on tcgen05, fp32 accumulate)",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "eager_evals_per_s": eager_evals}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stage_times(lib, F, torch, sets, B, K, T, dev, reps=40):
    """CUDA-event time of each stage of one eval, called through the C ABI, inputs cold in L2."""
    from kccotgan_b200 import _lib
    J = 8
    nsets = len(sets)
    C3 = torch.empty(3, B, B, device=dev)
    Cb = torch.empty_like(C3)
    ws = torch.empty(lib.kccot_mixed_cost_workspace_bytes(1, B, K), dtype=torch.uint8, device=dev)
    ws2 = torch.empty(max(256, lib.kccot_sinkhorn_workspace_bytes(3, B, 100)), dtype=torch.uint8, device=dev)
    ws3 = torch.empty(lib.kccot_mixed_cost_bwd_workspace_bytes(1, B, K), dtype=torch.uint8, device=dev)
    uh = torch.empty(3, 101, B, device=dev)
    vh = torch.empty_like(uh)
    nits = torch.empty(3, dtype=torch.int32, device=dev)
    cost = torch.empty(3, device=dev)
    g3 = torch.tensor([2.0, -1.0, -1.0], device=dev)
    gf = torch.empty(B, K, device=dev)
    gh = [torch.empty(B, T, J, device=dev) for _ in range(4)]
    st = F._stream(dev)
    p = F._ptr

    def views(i):
        r, f, hf, mr, hr, mf = [t.detach() for t in sets[i % nsets]]
        return r.reshape(B, -1), f.reshape(B, -1), hf, mr, hr, mf

    def f_partials(i):
        r, f, hf, mr, hr, mf = views(i)
        _lib.call("kccot_mixed_sqdist_partials", p(r), p(f), 1, B, K, p(ws), ws.numel(), 0, st)

    def f_cost(i):
        r, f, hf, mr, hr, mf = views(i)
        _lib.call("kccot_mixed_cost_fwd", p(r), p(f), 1, B, K, p(hf), p(mr), p(hr), p(mf), T, J, S, p(C3), p(ws),
                  ws.numel(), 0, st)

    def f_skf(i):
        _lib.call("kccot_sinkhorn_fwd", p(C3), 3, B, 1.0, 100, 100, 1e-2, 0, p(uh), p(vh), p(nits), p(cost), p(ws2),
                  ws2.numel(), st)

    def f_skb(i):
        _lib.call("kccot_sinkhorn_bwd", p(C3), 3, B, 1.0, 100, p(uh), p(vh), p(nits), p(g3), p(Cb), p(ws2),
                  ws2.numel(), st)

    def f_grad(i):
        r, f, hf, mr, hr, mf = views(i)
        _lib.call("kccot_mixed_cost_bwd", p(Cb), p(r), p(f), 1, B, K, p(hf), p(mr), p(hr), p(mf), T, J, S, None,
                  p(gf), p(gh[0]), p(gh[1]), p(gh[2]), p(gh[3]), p(ws3), ws3.numel(), 0, st)

    def f_grad_only(i):
        r, f, hf, mr, hr, mf = views(i)
        _lib.call("kccot_mixed_cost_bwd", p(Cb), p(r), p(f), 1, B, K, p(hf), p(mr), p(hr), p(mf), T, J, S, None,
                  p(gf), None, None, None, None, p(ws3), ws3.numel(), 0, st)

    out = {}
    for name, fn in (("sqdist_tc_kernel", f_partials), ("cost_fwd(sqdist+finalize)", f_cost),
                     ("sinkhorn_fwd_small_kernel", f_skf), ("sinkhorn_bwd_small_kernel", f_skb),
                     ("cost_bwd(W+grad+martingale)", f_grad), ("grad_tc_kernel", f_grad_only)):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        # back-to-back launches (input sets rotate, so the videos are cold in L2): launch latency overlaps
        # execution and the event pair brackets `reps` launches on the launching stream
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i + 3)
        b.record()
        torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / reps * 1e3
    # grad_only still includes the tiny W-build kernel (~2 us); the ncu launch list separates them
    return out


if __name__ == "__main__":
    main()
