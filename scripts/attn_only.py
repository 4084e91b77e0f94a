"""Runs only the MID-FC attention module forward (+ backward with --bwd): target for ncu captures."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from csn_b200 import midfc, synth

h = 1
bwd = "--bwd" in sys.argv
m = midfc.MultiHeadAttention(h, 256, 256, 256).cuda().eval()
x = torch.relu(torch.randn(8, 256, 10000, 1, device="cuda"))
y = torch.relu(torch.randn(8, 256, 10000, 1, device="cuda"))
for _ in range(3):
    if bwd:
        out, _ = m(x, y, y, "test")
        out.square().mean().backward()
    else:
        with torch.no_grad():
            out, _ = m(x, y, y, "test")
torch.cuda.synchronize()
print("ok", float(out.sum()))
