import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from csn_b200 import midfc, synth

def run(fused, h=2, B=2):
    os.environ["CSN_FUSED_ATTN"] = "1" if fused else "0"
    m = midfc.MultiHeadAttention(h, 256, 256, 256).cuda().eval()
    sd = synth.midfc_state(3, h)
    m.load_state_dict({k[len("attention."):]: v for k, v in sd.items() if k.startswith("attention.")})
    g = synth.gen(4)
    xq = synth.iid_features(g, B).cuda()
    xkv = synth.iid_features(g, B).cuda()
    outs = []
    with torch.no_grad():
        for _ in range(4):
            y, _ = m(xq, xkv, xkv, "test")
            outs.append(y.clone())
    for i in range(1, 4):
        d = (outs[i] - outs[0]).abs()
        print(f"fused={fused} run{i} vs run0: max diff {d.max().item():.3e}, n_diff {(d > 0).sum().item()}, "
              f"rows_diff {(d.amax(-1) > 0).sum().item()}")
        if d.max() > 0:
            idx = (d.amax(-1) > 0).nonzero()[:10].tolist()
            print("   first differing (batch,row):", idx)
run(False); run(True)

def run_csa():
    os.environ.pop("CSN_FUSED_ATTN", None)
    m = midfc.get_model("csa", 15, 2, 2).cuda().eval()
    m.load_state_dict(synth.midfc_state(22, 2, 15))
    x, nb = synth.csa_batch(23, 2, 2)
    x = x.cuda(); nbc = nb.cuda()
    with torch.no_grad():
        a = m.get_csa_feats(x, nbc, "test"); b = m.get_csa_feats(x, nbc, "test"); c = m.get_csa_feats(x, nb, "test")
        la, lb = m.logit(a), m.logit(b)
    print("csa feats run-to-run max diff", (a - b).abs().max().item(), "cpu-neighbours vs gpu", (a - c).abs().max().item(),
          "logits", (la - lb).abs().max().item())
run_csa()
