"""Development probe: throughput of cfg2 evaluations replayed on 1 / 2 / 3 streams (independent evaluations overlap)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kccotgan_b200.graphed import GraphedSinkhornLoss  # noqa: E402
from kccotgan_b200.synthetic import CONFIGS, INPUT_ORDER, make_inputs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_mazes"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
dev = torch.device("cuda", 0)
cfg = {k: v for k, v in CONFIGS[name].items() if k != "nprob"}
nsets = 12
graphs = []
for i in range(nsets):
    inp = make_inputs(J=8, kind="uniform", seed=1 + 1000 * i, device=dev, **cfg)
    graphs.append(GraphedSinkhornLoss(*[inp[k] for k in INPUT_ORDER], 1.0 / 15.0, adopt=True))
torch.cuda.synchronize()
ref = []
for g in graphs:
    g.step()
    torch.cuda.synchronize()
    ref.append((float(g.loss), g.grads["fake"].clone()))
for ns in (1, 4, 6, 8, 12):
    streams = [torch.cuda.Stream(dev) for _ in range(ns)]
    def run(n):
        for i in range(n):
            with torch.cuda.stream(streams[i % ns]):
                graphs[i % nsets].graph.replay()
    run(60)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    e0.record(main)
    for s in streams:
        s.wait_stream(main)
    run(steps)
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = all(float(g.loss) == r[0] and torch.equal(g.grads["fake"], r[1]) for g, r in zip(graphs, ref))
    print(f"{name}: {ns} stream(s): {steps / ms * 1e3:9.0f} evals/s  ({ms / steps * 1e3:7.2f} us per eval)  results identical: {ok}", flush=True)
