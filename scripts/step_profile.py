"""Per-kernel device time of one CSA training step (torch.profiler / CUPTI), config 2 of BASELINE.json."""
import sys, collections
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
from csn_b200 import midfc, synth
import bench

train = "--train" in sys.argv   # model.train(): dropout on
argv = [a for a in sys.argv[1:] if a != "--train"]
h = int(argv[0]) if argv else 1
dev = torch.device("cuda:0")
model = midfc.get_model("csa", 15, h, 3).to(dev)
model = model.train() if train else model.eval()
model.load_state_dict(synth.midfc_state(1, h, 15))
g = torch.Generator(device=dev).manual_seed(1)
nb = torch.relu(torch.randn(8, 4, 256, 10000, 1, device=dev, generator=g))
x = nb[:, 0].clone()
lab = torch.randint(0, 15, (8, 10000), device=dev, generator=g)
params = [p for n, p in model.named_parameters() if not n.startswith("fc_1")]
def step():
    for p in params: p.grad = None
    loss = model.forward_loss(x, "test", nb, lab); loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
n = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n): step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
per = len(evs) // n
print("--- kernels of the last step in launch order (>= 20 us)")
for e in evs[-per:]:
    if e.device_time >= 2:
        print(f"{e.device_time:9.1f} us  {e.name[:90]}")
tot = collections.OrderedDict()
for e in prof.events():
    if e.device_type.name != "CUDA": continue
    k = e.name[:70]
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += e.device_time
allt = sum(v for _, v in tot.values())
print(f"sum of kernel time per step: {allt / n:.1f} us")
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f"{v / n:9.1f} us  n={c / n:5.1f}  {k}")
