"""Development probe: which kernels make up one training iteration with the stub networks (torch profiler)."""
import os
import sys
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kccotgan_b200.synthetic import CONFIGS  # noqa: E402
from kccotgan_b200.train_step import StubDiscriminator, StubGenerator, make_training_steps  # noqa: E402
c = CONFIGS["cfg2_mazes"]
B, T, ctx, H, W, C = c["B"], c["T"], c["ctx"], c["H"], c["W"], c["C"]
dev = torch.device("cuda")
torch.manual_seed(1)
gen = StubGenerator(T - ctx, C).to(dev)
dh, dm = StubDiscriminator(H, W, C).to(dev), StubDiscriminator(H, W, C).to(dev)
disc_step, gen_step = make_training_steps(gen, dh, dm, B)
x = torch.rand(B, H, T, W, C, device=dev)
for _ in range(3):
    disc_step(x[:, :, :ctx], x[:, :, ctx:], 5.0); gen_step(x[:, :, :ctx], x[:, :, ctx:], 5.0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        disc_step(x[:, :, :ctx], x[:, :, ctx:], 5.0); gen_step(x[:, :, :ctx], x[:, :, ctx:], 5.0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
