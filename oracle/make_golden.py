"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on the seeded synthetic inputs of csn_b200/synth.py.  Run in the build container only
(`python oracle/make_golden.py`); the GPU box has no /root/reference and only reads the committed
vectors.  Outputs are sub-sampled (strided) to keep fixtures small; sums / sums of squares of the
full tensors are stored next to the samples so that a wrong element outside the sample still shows.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from csn_b200 import synth  # noqa: E402

REF = Path("/root/reference")
GOLDEN = ROOT / "tests" / "golden"


def load_reference():
    """(csa_models, mink_attention) — import recipes from SURVEY.md Appendix C."""
    sys.path.insert(0, str(REF / "MID-FC"))
    import csa_models as ref_midfc  # type: ignore

    me = types.ModuleType("MinkowskiEngine")
    me.SparseTensor = type("SparseTensor", (), {})
    sys.modules.setdefault("MinkowskiEngine", me)
    spec = importlib.util.spec_from_file_location("ref_mink_attention", REF / "MinkowskiNet/models/attention.py")
    ref_mink = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mink)
    return ref_midfc, ref_mink


def sample(t: torch.Tensor, max_elems: int = 4096) -> dict[str, np.ndarray]:
    """Strided sample + checksums of a tensor (see tests/golden_util.py for the reader)."""
    flat = t.detach().reshape(-1).double()
    n = flat.numel()
    stride = max(1, n // max_elems)
    if stride > 1 and stride % 2 == 0:
        stride += 1  # odd stride: walks across rows and columns of power-of-two shapes
    idx = torch.arange(0, n, stride)
    return {
        "shape": np.array(t.shape, dtype=np.int64),
        "stride": np.array(stride, dtype=np.int64),
        "values": flat[idx].numpy().astype(np.float64 if t.dtype == torch.float64 else np.float32),
        "sum": np.array(flat.sum().item()),
        "sumsq": np.array((flat * flat).sum().item()),
    }


def pack(prefix: str, d: dict[str, np.ndarray]) -> dict[str, np.ndarray]:
    return {f"{prefix}.{k}": v for k, v in d.items()}


def labels_for(seed: int, batch: int, n: int, num_classes: int) -> torch.Tensor:
    return torch.randint(0, num_classes, (batch, n), generator=synth.gen(seed))


def ref_masked_ce(logits, label, num_classes):
    """The reference's training loss (MID-FC/csa_training.py:94-108) is defined inside a script
    that parses sys.argv at import; the three lines are applied here with torch's own CE."""
    lg = logits.squeeze(-1).permute(0, 2, 1).contiguous().view(-1, num_classes)
    lb = label.view(-1)
    keep = torch.where(lb > 0)[0]
    return torch.nn.CrossEntropyLoss()(lg[keep], lb[keep])


def golden_midfc_mha(ref_midfc, seed=11, n_heads=1):
    sd = synth.midfc_state(seed, n_heads)
    m = ref_midfc.MultiHeadAttention(n_heads, 256, 256, 256).eval()
    m.load_state_dict({k[len("attention."):]: v for k, v in sd.items() if k.startswith("attention.")})
    g = synth.gen(seed + 1)
    xq = synth.iid_features(g, 1)
    xkv = synth.iid_features(g, 1)
    with torch.no_grad():
        y, attn = m(xq, xkv, xkv, "test")
    out = {"seed": np.array(seed), "n_heads": np.array(n_heads)}
    out.update(pack("y", sample(y)))
    out.update(pack("attn", sample(attn)))
    return out


def _fp64_compat_grads(ref_midfc, sd, n_heads, K, x, nb, label, num_classes, grads32):
    """The same step in fp64 (the reference module cast with .double()): the compatibility_{q,k} gradients are a
    difference of nearly equal numbers, so fp32 itself is only good to 1e-4..1e-2 there; tests gate the kernels
    against THESE values with the reference's own fp32 error as the yardstick."""
    m64 = ref_midfc.get_model("csa", num_classes, n_heads, K).eval()
    m64.load_state_dict(sd)
    m64 = m64.double()
    x64 = x.detach().double().requires_grad_(True)
    logits = m64.logit(m64.get_csa_feats(x64, nb.double(), "test"))
    ref_masked_ce(logits, label, num_classes).backward()
    out = {}
    for name, p in m64.named_parameters():
        if p.grad is None or not name.startswith("compatibility"):
            continue
        out.update(pack("grad64." + name, sample(p.grad)))
        g32 = grads32[name].double()
        out[f"grad64.{name}.ref32_rel_err"] = np.array(float((g32 - p.grad).norm() / p.grad.norm()))
    return out


def golden_midfc_csa(ref_midfc, seed, n_heads, K, batch, num_classes=15, max_elems=4096):
    sd = synth.midfc_state(seed, n_heads, num_classes)
    m = ref_midfc.get_model("csa", num_classes, n_heads, K).eval()
    m.load_state_dict(sd)
    ref_midfc.device = torch.device("cpu")
    x, nb = synth.csa_batch(seed + 1, batch, K)
    label = labels_for(seed + 2, batch, x.shape[2], num_classes)
    x = x.clone().requires_grad_(True)
    feats = m.get_csa_feats(x, nb, "test")
    logits = m.logit(feats)
    loss = ref_masked_ce(logits, label, num_classes)
    loss.backward()
    out = {"seed": np.array(seed), "n_heads": np.array(n_heads), "K": np.array(K), "batch": np.array(batch),
           "num_classes": np.array(num_classes), "loss": np.array(loss.item())}
    out.update(pack("feats", sample(feats, max_elems)))
    out.update(pack("logits", sample(logits, max_elems)))
    out.update(pack("grad.x", sample(x.grad, max_elems)))
    grads32 = {}
    for name, p in m.named_parameters():
        if p.grad is not None:
            out.update(pack("grad." + name, sample(p.grad, max_elems)))
            grads32[name] = p.grad.detach().clone()
    with torch.no_grad():
        ssa, _ = m.get_ssa_feats(x.detach(), "test")
    out.update(pack("ssa", sample(ssa, max_elems)))
    out.update(_fp64_compat_grads(ref_midfc, sd, n_heads, K, x, nb, label, num_classes, grads32))
    return out


def golden_midfc_mha_prefix(ref_midfc, seed=13, n_heads=1, n_used=5000):
    """Config-5 parity point (N = 5000, iters = 10): the reference hard-codes 20 chunks of 500 points
    (csa_models.py:83-84), but its attention is block-diagonal, so the first n_used rows of its output on a
    10 000-point input ARE the iters = n_used/500 result on the first n_used points; the loss touches only them."""
    sd = synth.midfc_state(seed, n_heads)
    m = ref_midfc.MultiHeadAttention(n_heads, 256, 256, 256).eval()
    m.load_state_dict({k[len("attention."):]: v for k, v in sd.items() if k.startswith("attention.")})
    g = synth.gen(seed + 1)
    xq = synth.iid_features(g, 1)
    xkv = synth.iid_features(g, 1)
    gy = torch.randn(1, n_used, 256, generator=g)
    y, _ = m(xq, xkv, xkv, "test")
    (y[:, :n_used] * gy).sum().backward()
    out = {"seed": np.array(seed), "n_heads": np.array(n_heads), "n_used": np.array(n_used)}
    out.update(pack("y", sample(y[:, :n_used].contiguous())))
    for name, p in m.named_parameters():
        out.update(pack("grad." + name, sample(p.grad)))
    return out


def golden_knn_graph(ref_midfc, seed, n_shapes, n_points, K, n_categories):
    """A collection large enough that top-(K+1) has to discriminate among hundreds of candidates."""
    m = ref_midfc.get_model("csa", 15, 1, K).eval()
    f = synth.clustered_shapes(seed, n_shapes, n_points=n_points, n_categories=n_categories)
    with torch.no_grad():
        scores = m.get_retrieval_measure(f, f)
    return {"seed": np.array(seed), "n_shapes": np.array(n_shapes), "n_points": np.array(n_points),
            "K": np.array(K), "n_categories": np.array(n_categories),
            "scores": scores.numpy(), "graph": scores.topk(K + 1, dim=-1).indices.numpy()}


def golden_mink_csa_head(ref_mink, seed=51, n_head=4, K=2, lens=(260, 140), key_lens=((190, 310), (120, 90))):
    """CSA block of HRNetSimCSN.forward (hrnet.py:370-417) with the reference's OWN MultiHeadAttention module
    (attention.py, imported) and the block's glue applied line by line on dense per-shape tensors (the rest of
    hrnet.py needs MinkowskiEngine): forward features, and the gradients of a seeded linear loss w.r.t. every
    parameter and every input (the backbone is trained end to end, trainer_csn.py:200-210)."""
    import math
    sd = synth.mink_state(seed, n_head)
    mha = ref_mink.MultiHeadAttention(n_head, 256, 256 // n_head, 256 // n_head).eval()
    mha.load_state_dict({k[len("MHA."):]: v for k, v in sd.items() if k.startswith("MHA.")})
    lin_q = torch.nn.Linear(256, 256, bias=False)
    lin_k = torch.nn.Linear(256, 256, bias=False)
    lin_q.weight.data.copy_(sd["linear_q.weight"]); lin_k.weight.data.copy_(sd["linear_k.weight"])
    g = synth.gen(seed + 1)
    q_feats = [torch.relu(torch.randn(L, 256, generator=g)).requires_grad_(True) for L in lens]
    k_feats = [[torch.relu(torch.randn(L, 256, generator=g)).requires_grad_(True) for L in kl] for kl in key_lens]   # [K][B]
    gys = [torch.randn(L, 256, generator=g) for L in lens]

    def ssa(f):   # hrnet.py:456-470
        return mha(f[None], f[None], f[None])[0][0]

    q_ssa = [ssa(f) for f in q_feats]
    key_ssa = [q_ssa] + [[ssa(f) for f in kf] for kf in k_feats]
    outs = []
    loss = 0.0
    for b in range(len(lens)):   # hrnet.py:376-417
        gq = torch.nn.functional.normalize(lin_q(q_ssa[b].mean(dim=0)), dim=-1)
        sims = []
        for ks in key_ssa:
            gk = torch.nn.functional.normalize(lin_k(ks[b].mean(dim=0)), dim=-1)
            sims.append((gq * gk).sum() / math.sqrt(256.0))   # ScaledDotProduct(temperature=sqrt(d_model)), :357
        comp = torch.softmax(torch.stack(sims), dim=0)
        csa = comp[0] * q_ssa[b]
        for i in range(K):
            cross = mha(q_feats[b][None], k_feats[i][b][None], k_feats[i][b][None])[0][0]
            csa = csa + comp[i + 1] * cross
        outs.append(csa)
        loss = loss + (csa * gys[b]).sum()
    loss.backward()
    res = {"seed": np.array(seed), "n_head": np.array(n_head), "K": np.array(K), "lens": np.array(lens),
           "key_lens": np.array(key_lens)}
    for b, o in enumerate(outs):
        res.update(pack(f"out{b}", sample(o)))
        res.update(pack(f"grad.q{b}", sample(q_feats[b].grad)))
        for i in range(K):
            res.update(pack(f"grad.k{i}_{b}", sample(k_feats[i][b].grad)))
    for name, p in mha.named_parameters():
        res.update(pack("grad.MHA." + name, sample(p.grad)))
    res.update(pack("grad.linear_q.weight", sample(lin_q.weight.grad)))
    res.update(pack("grad.linear_k.weight", sample(lin_k.weight.grad)))
    return res


def golden_knn(ref_midfc, seed, n_shapes, n_points, K, n_categories):
    m = ref_midfc.get_model("csa", 15, 1, K).eval()
    f = synth.clustered_shapes(seed, n_shapes, n_points=n_points, n_categories=n_categories)
    with torch.no_grad():
        scores = m.get_retrieval_measure(f, f)
        graph = m.get_knn_graph(f, f, K)
        # rectangular: first 3 shapes as queries against the rest as candidates
        scores_rect = m.get_retrieval_measure(f[:3], f[3:])
    return {"seed": np.array(seed), "n_shapes": np.array(n_shapes), "n_points": np.array(n_points),
            "K": np.array(K), "n_categories": np.array(n_categories),
            "scores": scores.numpy(), "graph": graph.numpy(), "scores_rect": scores_rect.numpy()}


def golden_mink_mha(ref_mink, seed=31, n_head=4, Lq=300, Lk=200):
    sd = synth.mink_state(seed, n_head)
    m = ref_mink.MultiHeadAttention(n_head, 256, 256 // n_head, 256 // n_head).eval()
    m.load_state_dict({k[len("MHA."):]: v for k, v in sd.items() if k.startswith("MHA.")})
    g = synth.gen(seed + 1)
    q = torch.relu(torch.randn(1, Lq, 256, generator=g)).requires_grad_(True)
    k = torch.relu(torch.randn(1, Lk, 256, generator=g)).requires_grad_(True)
    out, attn = m(q, k, k)
    gy = torch.randn(out.shape, generator=g)
    (out * gy).sum().backward()
    res = {"seed": np.array(seed), "n_head": np.array(n_head), "Lq": np.array(Lq), "Lk": np.array(Lk)}
    res.update(pack("out", sample(out)))
    res.update(pack("attn", sample(attn)))
    res.update(pack("grad.q", sample(q.grad)))
    res.update(pack("grad.k", sample(k.grad)))
    for name, p in m.named_parameters():
        res.update(pack("grad." + name, sample(p.grad)))
    # ScaledDotProduct (attention.py:78-108) on dense inputs
    sdp = ref_mink.ScaledDotProduct(16.0)
    with torch.no_grad():
        res.update(pack("sdp", sample(sdp(q.detach()[0, :5], k.detach()[0, :7]))))
    return res


def main() -> None:
    if not REF.exists():
        raise SystemExit("/root/reference is not present: golden vectors can only be regenerated in the build container")
    torch.set_num_threads(8)
    torch.manual_seed(0)
    ref_midfc, ref_mink = load_reference()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    jobs = {
        "midfc_mha_h1": lambda: golden_midfc_mha(ref_midfc, seed=11, n_heads=1),
        "midfc_mha_h2": lambda: golden_midfc_mha(ref_midfc, seed=12, n_heads=2),
        "midfc_csa_cfg1": lambda: golden_midfc_csa(ref_midfc, seed=21, n_heads=1, K=1, batch=1),
        "midfc_csa_b2_k2_h2": lambda: golden_midfc_csa(ref_midfc, seed=22, n_heads=2, K=2, batch=2),
        "midfc_csa_b8_k3_h1": lambda: golden_midfc_csa(ref_midfc, seed=23, n_heads=1, K=3, batch=8, max_elems=2048),
        "midfc_mha_n5000": lambda: golden_midfc_mha_prefix(ref_midfc),
        "knn_graph_s200": lambda: golden_knn_graph(ref_midfc, seed=43, n_shapes=200, n_points=1000, K=4, n_categories=8),
        "mink_csa_head": lambda: golden_mink_csa_head(ref_mink),
        "knn_small": lambda: golden_knn(ref_midfc, seed=41, n_shapes=12, n_points=1000, K=3, n_categories=4),
        "knn_10k": lambda: golden_knn(ref_midfc, seed=42, n_shapes=6, n_points=10000, K=2, n_categories=3),
        "mink_mha": lambda: golden_mink_mha(ref_mink),
    }
    only = set(sys.argv[1:])
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        data = fn()
        np.savez_compressed(GOLDEN / f"{name}.npz", **data)
        print(f"wrote {name}.npz ({(GOLDEN / (name + '.npz')).stat().st_size / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
