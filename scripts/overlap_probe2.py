"""Development probe: is the lanes throughput launch-bound?  Host time of the submit loop against the device time;
one fused graph of all lanes against separate graphs."""
import os
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kccotgan_b200.graphed import EvaluationLanes, GraphedSinkhornLoss  # noqa: E402
from kccotgan_b200.synthetic import CONFIGS, INPUT_ORDER, make_inputs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_mazes"
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
ctxf = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = 1200
dev = torch.device("cuda", 0)
cfg = {k: v for k, v in CONFIGS[name].items() if k != "nprob"}
nl = 6
evs = []
for i in range(nl):
    inp = make_inputs(J=8, kind=kind, seed=1 + 1000 * i, device=dev, **cfg)
    evs.append(GraphedSinkhornLoss(*[inp[k] for k in INPUT_ORDER], 1.0 / 15.0, adopt=True, ctx_frames=ctxf))
lanes = EvaluationLanes(evs, nl)
torch.cuda.synchronize()

def run(n):
    lanes.fork()
    for i in range(n):
        lanes.submit(i % nl)
    lanes.join()

run(60)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
run(steps)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"{name} {kind} ctx={ctxf}: separate graphs: device {e0.elapsed_time(e1) / steps * 1e3:.2f} us/eval, host submit "
      f"{(t1 - t0) / steps * 1e6:.2f} us/eval, host until done {(t2 - t0) / steps * 1e6:.2f} us/eval", flush=True)

# one fused graph: all lanes as parallel branches
g = torch.cuda.CUDAGraph()
cap = torch.cuda.Stream(dev)
side = [torch.cuda.Stream(dev) for _ in range(nl)]
torch.cuda.synchronize()
with torch.cuda.graph(g, stream=cap):
    cur = torch.cuda.current_stream(dev)
    outs = []
    for s, ev in zip(side, evs):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(ev._eager())
    for s in side:
        cur.wait_stream(s)
torch.cuda.synchronize()
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record()
for _ in range(steps // nl):
    g.replay()
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
n = steps // nl * nl
print(f"   fused graph of {nl} lanes: device {e0.elapsed_time(e1) / n * 1e3:.2f} us/eval, host submit "
      f"{(t1 - t0) / n * 1e6:.2f} us/eval", flush=True)
ok = all(float(o[0]) == float(ev.loss) and torch.equal(o[1][0], ev.grads["fake"]) for o, ev in zip(outs, evs))
print("   fused results identical to the separate graphs:", ok, flush=True)
