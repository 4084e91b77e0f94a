"""One MinkowskiNet CSA-head step for an ncu capture of the d_head-64 attention kernels (run under ncu with
-k regex:attn_ ; the first step is the only one)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from csn_b200 import mink, synth

dev = torch.device("cuda:0")
head = mink.CSAHead(256, 4, precision="bf16").to(dev).eval()
head.load_state_dict(synth.mink_state(2, 4), strict=False)
B, K = 8, 3
lens = synth.ragged_lengths(7, B * (K + 1))
g = synth.gen(8)
q = [torch.relu(torch.randn(lens[b], 256, generator=g)).to(dev).requires_grad_(True) for b in range(B)]
keys = [[torch.relu(torch.randn(lens[B * (k + 1) + b], 256, generator=g)).to(dev) for b in range(B)] for k in range(K)]
out = head(q, keys)
sum(o.square().mean() for o in out).backward()
torch.cuda.synchronize()
print("ok")
