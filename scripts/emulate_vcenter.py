"""CPU experiment (no GPU needed): how much of the compatibility-gradient error comes from 16-bit O / P / V, and what
centring V on its per-chunk key mean buys.  The forward pass is emulated in fp64 with fp16 rounding applied at the
places the kernels round (straight-through for autograd), and the compatibility_{q,k} gradients are compared with
the reference's fp64 values (tests/golden/*.npz, grad64.*)."""
import math
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from csn_b200 import synth  # noqa: E402


def q16(x, on=True):
    if not on:
        return x
    return x + (x.detach().to(torch.float16).to(x.dtype) - x.detach())


def mha(xq, xkv, w, center, quant, olo=False):
    """xq, xkv: (N, 256) fp64 rows. Returns LN output (N,256)."""
    wq, wk, wv, wo = (w[f"attention.{n}.weight"] for n in ("w_qs", "w_ks", "w_vs", "fc"))
    Xq16, Xk16 = q16(xq, quant), q16(xkv, quant)
    Q = q16(Xq16 @ q16(wq, quant).t(), quant)
    K = q16(Xk16 @ q16(wk, quant).t(), quant)
    Vf = Xk16 @ q16(wv, quant).t()
    N = xq.shape[0]
    Qc, Kc, Vc = Q.view(20, 500, 256), K.view(20, 500, 256), Vf.view(20, 500, 256)
    c = Vc.mean(dim=1, keepdim=True).detach() if center else torch.zeros(20, 1, 256, dtype=Vf.dtype)
    V16 = q16(Vc - c, quant)
    S = (Qc @ Kc.transpose(1, 2)) / 16.0
    m = S.max(dim=-1, keepdim=True).values.detach()
    p = torch.exp(S - m)
    l = p.sum(-1, keepdim=True)
    O = (q16(p, quant) @ V16) / l
    O16 = q16(O, quant)
    if olo and quant:   # hi + lo residual (22-bit) as used for the colsum blocks
        O16 = O16 + q16((O - O16.detach()) * 2048.0, True) / 2048.0
    Z = (O16 @ q16(wo, quant).t().unsqueeze(0)) + (c @ wo.t()) + xq.view(20, 500, 256)
    Z = Z.reshape(N, 256)
    return F.layer_norm(Z, (256,), w["attention.norm.weight"], w["attention.norm.bias"], 1e-6)


def run(name, center, quant, olo=False):
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    seed, h, K, B, C = (int(g[k]) for k in ("seed", "n_heads", "K", "batch", "num_classes"))
    assert h == 1 and B == 1
    w = {k: v.double().clone().requires_grad_(v.is_floating_point()) for k, v in synth.midfc_state(seed, h, C).items()}
    x, nb = synth.csa_batch(seed + 1, B, K)
    label = torch.randint(0, C, (B, x.shape[2]), generator=synth.gen(seed + 2))
    X = [x[0, :, :, 0].t().double()] + [nb[0, k, :, :, 0].t().double() for k in range(1, K + 1)]
    ssa = [mha(X[k], X[k], w, center, quant, olo) for k in range(K + 1)]
    pooled = torch.stack([s.mean(0) for s in ssa])
    uq = F.normalize(pooled[0:1] @ w["compatibility_q.weight"].t() + w["compatibility_q.bias"], dim=-1)
    uk = F.normalize(pooled @ w["compatibility_k.weight"].t() + w["compatibility_k.bias"], dim=-1)
    comp = torch.softmax((uq * uk).sum(-1), dim=0)
    out = comp[0] * ssa[0]
    for k in range(1, K + 1):
        out = out + comp[k] * mha(X[0], X[k], w, center, quant, olo)
    logits = out @ w["logit.weight"].view(C, 256).t()
    keep = label[0] > 0
    loss = F.cross_entropy(logits[keep], label[0][keep])
    loss.backward()
    res = {}
    for pname in ("compatibility_q.weight", "compatibility_q.bias", "compatibility_k.weight", "compatibility_k.bias",
                  "attention.w_qs.weight", "attention.fc.weight"):
        key = f"grad64.{pname}" if pname.startswith("compat") else f"grad.{pname}"
        stride = int(g[f"{key}.stride"])
        got = w[pname].grad.reshape(-1)[::stride]
        want = torch.from_numpy(g[f"{key}.values"].astype("float64"))
        res[pname] = float((got - want).norm() / want.norm())
    return loss.item(), res


if __name__ == "__main__":
    torch.set_num_threads(8)
    name = "midfc_csa_cfg1"
    g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    print("reference fp32 vs fp64:", {k[7:-14]: float(g[k]) for k in g.files if k.endswith("ref32_rel_err")})
    for center, quant, olo in ((False, False, False), (False, True, False), (False, True, True), (True, True, False)):
        loss, res = run(name, center, quant, olo)
        print(f"center={center} quant={quant} olo={olo} loss={loss:.6f}", {k: f"{v:.2e}" for k, v in res.items()})
