"""Quick throughput probe of csn_knn_scores on synthetic clustered shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from csn_b200 import knn, synth

nq, nc, N = (int(a) for a in (sys.argv[1:4] + ["8", "64", "10000"][len(sys.argv) - 1:]))
f = synth.clustered_shapes(1, max(nq, nc), n_points=N, n_categories=4).cuda()
q = knn.build_store(f[:nq].contiguous())
c = knn.build_store(f[:nc].contiguous())
for _ in range(2):
    s = knn.scores_from_stores(q, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 3
e0.record()
for _ in range(iters):
    s = knn.scores_from_stores(q, c)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 2.0 * nq * nc * N * N * 256
print(f"knn {nq}x{nc} shapes N={N}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s  ({nq * nc / ms * 1e3:.0f} pairs/s)")
