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


def golden_midfc_csa(ref_midfc, seed, n_heads, K, batch, num_classes=15):
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
    out.update(pack("feats", sample(feats)))
    out.update(pack("logits", sample(logits)))
    out.update(pack("grad.x", sample(x.grad)))
    for name, p in m.named_parameters():
        if p.grad is not None:
            out.update(pack("grad." + name, sample(p.grad)))
    with torch.no_grad():
        ssa, _ = m.get_ssa_feats(x.detach(), "test")
    out.update(pack("ssa", sample(ssa)))
    return out


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
