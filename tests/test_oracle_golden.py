"""CPU: the oracle restatement (oracle/csa_oracle.py) against vectors produced by the reference
itself (oracle/make_golden.py).  This is what pins the oracle (the reference ships
no tests or fixtures of its own)."""
import numpy as np
import pytest
import torch

from csn_b200 import synth
from oracle import csa_oracle as O
from tests import golden_util as G

TOL = 2e-5  # fp32 vs fp32, different GEMM blocking / summation order


@pytest.mark.parametrize("name", ["midfc_mha_h1", "midfc_mha_h2"])
def test_midfc_mha(name):
    g = G.load(name)
    seed, h = int(g["seed"]), int(g["n_heads"])
    w = synth.midfc_state(seed, h)
    gen = synth.gen(seed + 1)
    xq = synth.iid_features(gen, 1)
    xkv = synth.iid_features(gen, 1)
    y, attn = O.mha_midfc(xq, xkv, xkv, w, h)
    G.compare_sampled(g, "y", y, TOL)
    G.compare_sampled(g, "attn", attn, TOL)


def test_midfc_mha_needs_10000_points():
    w = synth.midfc_state(1, 1)
    x = synth.iid_features(synth.gen(0), 1, n_points=2000)
    with pytest.raises(IndexError):
        O.mha_midfc(x, x, x, w, 1)


@pytest.mark.parametrize("name", ["midfc_csa_cfg1", "midfc_csa_b2_k2_h2", "midfc_csa_b8_k3_h1"])
def test_midfc_csa_forward_backward(name):
    g = G.load(name)
    seed, h, K, B, C = (int(g[k]) for k in ("seed", "n_heads", "K", "batch", "num_classes"))
    w = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in synth.midfc_state(seed, h, C).items()}
    x, nb = synth.csa_batch(seed + 1, B, K)
    label = torch.randint(0, C, (B, x.shape[2]), generator=synth.gen(seed + 2))
    x = x.clone().requires_grad_(True)
    feats = O.csa_feats(x, nb, w, h)
    logits = O.logits_1x1(feats, w)
    loss = O.masked_cross_entropy(logits, label)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    G.compare_sampled(g, "feats", feats, TOL)
    G.compare_sampled(g, "logits", logits, TOL)
    G.compare_sampled(g, "grad.x", x.grad, 1e-4)
    for key in g.files:
        if key.startswith("grad.") and key.endswith(".values") and key != "grad.x.values":
            pname = key[len("grad."):-len(".values")]
            # compatibility_* gradients are ill-conditioned in fp32 (the reference's own
            # fp32 values are 3e-4..3e-3 away from fp64); everything else agrees to ~1e-6.
            tol = 2e-2 if pname.startswith("compatibility") else 2e-4
            G.compare_sampled(g, "grad." + pname, w[pname].grad, tol, what=pname)
    with torch.no_grad():
        ssa, _ = O.ssa_feats(x.detach(), w, h)
    G.compare_sampled(g, "ssa", ssa, TOL)


@pytest.mark.parametrize("name", ["knn_small", "knn_10k"])
def test_knn(name):
    g = G.load(name)
    f = synth.clustered_shapes(int(g["seed"]), int(g["n_shapes"]), n_points=int(g["n_points"]),
                               n_categories=int(g["n_categories"]))
    scores = O.retrieval_measure(f, f)
    assert np.abs(scores.numpy() - g["scores"]).max() < 2e-6
    graph = O.knn_graph(f, f, int(g["K"]))
    assert np.array_equal(graph.numpy(), g["graph"])
    rect = O.retrieval_measure(f[:3], f[3:])
    assert np.abs(rect.numpy() - g["scores_rect"]).max() < 2e-6


def test_mink_mha():
    g = G.load("mink_mha")
    seed, h, Lq, Lk = (int(g[k]) for k in ("seed", "n_head", "Lq", "Lk"))
    w = {k: v.clone().requires_grad_(True) for k, v in synth.mink_state(seed, h).items()}
    gen = synth.gen(seed + 1)
    q = torch.relu(torch.randn(1, Lq, 256, generator=gen)).requires_grad_(True)
    k = torch.relu(torch.randn(1, Lk, 256, generator=gen)).requires_grad_(True)
    out, attn = O.mha_mink(q, k, k, w, h)
    gy = torch.randn(out.shape, generator=gen)
    (out * gy).sum().backward()
    G.compare_sampled(g, "out", out, TOL)
    G.compare_sampled(g, "attn", attn, TOL)
    G.compare_sampled(g, "grad.q", q.grad, 1e-4)
    G.compare_sampled(g, "grad.k", k.grad, 1e-4)
    for pname in ("w_qs.weight", "w_ks.weight", "w_vs.weight", "fc.weight", "norm.weight", "norm.bias"):
        G.compare_sampled(g, "grad." + pname, w["MHA." + pname].grad, 2e-4, what=pname)
    sdp = (q.detach()[0, :5] @ k.detach()[0, :7].t()) / 16.0
    G.compare_sampled(g, "sdp", sdp[None], TOL)


def test_midfc_mha_prefix_is_the_short_shape_result():
    """Config-5 fixture: the oracle with iters = 10 on the first 5 000 points equals the first 5 000 rows of the
    reference's 10 000-point output (block-diagonal attention)."""
    g = G.load("midfc_mha_n5000")
    seed, h, n_used = int(g["seed"]), int(g["n_heads"]), int(g["n_used"])
    w = {k: v.clone().requires_grad_(True) for k, v in synth.midfc_state(seed, h).items() if k.startswith("attention.")}
    gen = synth.gen(seed + 1)
    xq = synth.iid_features(gen, 1)[:, :, :n_used]
    xkv = synth.iid_features(gen, 1)[:, :, :n_used]
    gy = torch.randn(1, n_used, 256, generator=gen)
    y, _ = O.mha_midfc(xq, xkv, xkv, w, h, iters=n_used // 500)
    (y * gy).sum().backward()
    G.compare_sampled(g, "y", y, TOL)
    for k, v in w.items():
        G.compare_sampled(g, "grad." + k[len("attention."):], v.grad, 2e-4, what=k)


def test_mink_csa_head_forward_backward():
    """oracle.mink_csa_block (the restated glue of hrnet.py:370-417) against the fixture built from the reference's
    own MultiHeadAttention module: features and every gradient."""
    g = G.load("mink_csa_head")
    seed, h, K = int(g["seed"]), int(g["n_head"]), int(g["K"])
    lens, key_lens = [int(v) for v in g["lens"]], [[int(v) for v in row] for row in g["key_lens"]]
    w = {k: v.clone().requires_grad_(True) for k, v in synth.mink_state(seed, h).items()}
    gen = synth.gen(seed + 1)
    qf = [torch.relu(torch.randn(L, 256, generator=gen)).requires_grad_(True) for L in lens]
    kf = [[torch.relu(torch.randn(L, 256, generator=gen)).requires_grad_(True) for L in kl] for kl in key_lens]
    gys = [torch.randn(L, 256, generator=gen) for L in lens]
    outs = O.mink_csa_block(qf, kf, w, h)
    sum((o * gy).sum() for o, gy in zip(outs, gys)).backward()
    for b, o in enumerate(outs):
        G.compare_sampled(g, f"out{b}", o, TOL)
        G.compare_sampled(g, f"grad.q{b}", qf[b].grad, 2e-4)
        for i in range(K):
            G.compare_sampled(g, f"grad.k{i}_{b}", kf[i][b].grad, 2e-4)
    for k, v in w.items():
        if f"grad.{k}.values" in g.files:
            G.compare_sampled(g, "grad." + k, v.grad, 2e-2 if k.startswith("linear_") else 2e-4, what=k)
