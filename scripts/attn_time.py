"""Times the fused attention forward kernel alone (config-2 geometry: 56 blocks) under CSN_ATTN_DEBUG variants."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from csn_b200 import engine as E, _lib as L

dev = torch.device("cuda:0")
geom = E.Geometry()
NP = geom.rows_pad
S, nblk, h, d = 32, 56, 1, 256
HD = h * d
groups = [E.Group(n_in=S, n_out=1, blk0=0, q0=0, q_si=1, q_so=0, k0=0, k_si=1, k_so=0, v0=0, v_si=1, v_so=0),
          E.Group(n_in=3, n_out=8, blk0=S, q0=0, q_si=0, q_so=4, k0=1, k_si=1, k_so=4, v0=1, v_si=1, v_so=4)]
QKV = (torch.randn(S * NP, 3 * HD, device=dev) * 0.5).half()
O = torch.empty(nblk * NP, HD, dtype=torch.float16, device=dev)
lse = torch.empty(nblk * h * NP, device=dev)
items = E.attn_items(groups, geom, h, d, dev)
def run():
    rc = L.lib().csn_attn_fwd(QKV[:, :HD].data_ptr(), QKV[:, HD:2*HD].data_ptr(), QKV[:, 2*HD:].data_ptr(), S * NP, S * NP, HD,
                              3 * HD, 3 * HD, 3 * HD, d, L.CSN_F16, items.data_ptr(), items.shape[0], O.data_ptr(), O.shape[0], HD,
                              lse.data_ptr(), None, int(os.environ.get("PAIRED", "1")), 0, 0.0, L.stream_ptr())
    L.check(rc, "attn")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"debug={os.environ.get('CSN_ATTN_DEBUG','0')} paired={os.environ.get('PAIRED','1')}: {e0.elapsed_time(e1)/10*1e3:.1f} us per launch ({items.shape[0]} items)")
