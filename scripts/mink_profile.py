"""Per-kernel device time of one MinkowskiNet CSA-head step (config 4, K=3), torch.profiler."""
import sys, collections
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
from csn_b200 import mink, synth

dev = torch.device("cuda:0")
head = mink.CSAHead(256, 4, precision="bf16").to(dev).eval()
head.load_state_dict(synth.mink_state(2, 4), strict=False)
B, K = 8, 3
lens = synth.ragged_lengths(7, B * (K + 1))
g = synth.gen(8)
q = [torch.relu(torch.randn(lens[b], 256, generator=g)).to(dev).requires_grad_(True) for b in range(B)]
keys = [[torch.relu(torch.randn(lens[B * (k + 1) + b], 256, generator=g)).to(dev) for b in range(B)] for k in range(K)]
def step():
    for p in head.parameters(): p.grad = None
    out = head(q, keys)
    sum(o.square().mean() for o in out).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
tot = collections.OrderedDict()
for e in prof.events():
    if e.device_type.name != "CUDA": continue
    k = e.name[:90]
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += e.device_time
allt = sum(v for _, v in tot.values())
print(f"sum of kernel time: {allt:.1f} us")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{t:9.1f} us  n={n:4d}  {k}")
