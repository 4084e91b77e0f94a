"""Training-mode dropout (csa_models.py:56,115,136,141; attention.py:28,51,67,72).

The kernels regenerate the dropout masks from (seed, row id, column) with a counter-based hash (csrc/ptx.cuh) instead
of storing them.  Bit-parity with torch's Philox stream is not definable, so parity is established the
other way round: the SAME masks are rebuilt here on the host from the documented hash, the reference arithmetic is run
in fp64 with those masks, and outputs and gradients must agree to the usual 1e-3 — which proves that the forward, dV,
dS and fc kernels all see one and the same mask — plus the statistical properties (drop rate, mean preservation).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from csn_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu

M32 = np.uint64(0xFFFFFFFF)


def _mix(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x


def keep_mask(seed: int, row_ids: np.ndarray, n_cols: int, p: float) -> torch.Tensor:
    """(len(row_ids), n_cols) float64 mask * scale, exactly as csrc/ptx.cuh::drop_pair / drop_keep_{lo,hi}."""
    thresh = int(p * 65536 + 0.5)
    rk = (row_ids.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(seed)) & M32
    pairs = np.arange((n_cols + 1) // 2, dtype=np.uint64)
    kk = (pairs * np.uint64(0x85EBCA77) + np.uint64(0x165667B1)) & M32
    h = _mix(rk[:, None] ^ kk[None, :])
    lo, hi = (h & np.uint64(0xFFFF)) >= thresh, (h >> np.uint64(16)) >= thresh
    keep = np.stack([lo, hi], axis=-1).reshape(len(row_ids), -1)[:, :n_cols]
    return torch.from_numpy(keep.astype(np.float64)) * (65536.0 / (65536.0 - thresh))


def _seeds(torch_seed: int):
    """The (attention, fc) seeds the module derives for the first call after torch.manual_seed(torch_seed)."""
    torch.manual_seed(torch_seed)
    s = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    return s & 0x7FFFFFFF, (s * 2654435761 + 97) & 0x7FFFFFFF


def _ref_mha(xq, xkv, w, h, iters, chunk, chunk_pad, p, seeds, block=0):
    """fp64 reference of MultiHeadAttention.forward in train mode with the kernels' masks.  xq, xkv: (N, 256)."""
    d = w["w_qs.weight"].shape[0] // h
    NP = iters * chunk_pad
    ys = []
    for c in range(iters):
        q_in, kv_in = xq[c * chunk:(c + 1) * chunk], xkv[c * chunk:(c + 1) * chunk]
        Lq, Lk = q_in.shape[0], kv_in.shape[0]
        Q = (q_in @ w["w_qs.weight"].t()).view(Lq, h, d).transpose(0, 1)
        K = (kv_in @ w["w_ks.weight"].t()).view(Lk, h, d).transpose(0, 1)
        V = (kv_in @ w["w_vs.weight"].t()).view(Lk, h, d).transpose(0, 1)
        P = torch.softmax((Q / d ** 0.5) @ K.transpose(1, 2), dim=-1)
        masks = torch.stack([keep_mask(seeds[0], (block * h + a) * NP + c * chunk_pad + np.arange(Lq), Lk, p) for a in range(h)])
        O = ((P * masks) @ V).transpose(0, 1).reshape(Lq, h * d)
        fc = O @ w["fc.weight"].t()
        fc = fc * keep_mask(seeds[1], block * NP + c * chunk_pad + np.arange(Lq), 256, p)
        ys.append(F.layer_norm(fc + q_in, (256,), w["norm.weight"], w["norm.bias"], 1e-6))
    return torch.cat(ys, dim=0)


def test_midfc_mha_train_mode_matches_reference_with_the_same_masks():
    from csn_b200 import midfc
    h, seed = 1, 11
    sd = {k[len("attention."):]: v for k, v in synth.midfc_state(seed, h).items() if k.startswith("attention.")}
    m = midfc.MultiHeadAttention(h, 256, 256, 256).cuda().train()
    m.load_state_dict(sd)
    gen = synth.gen(seed + 1)
    xq, xkv = synth.iid_features(gen, 1), synth.iid_features(gen, 1)
    gy = torch.randn(1, 10000, 256, generator=gen)
    seeds = _seeds(123)
    torch.manual_seed(123)
    q_dev = xq.cuda().requires_grad_(True)
    kv_dev = xkv.cuda().requires_grad_(True)
    y, _ = m(q_dev, kv_dev, kv_dev, "train")
    (y * gy.cuda()).sum().backward()
    w = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xq64 = xq[0, :, :, 0].t().double().requires_grad_(True)
    xkv64 = xkv[0, :, :, 0].t().double().requires_grad_(True)
    want = _ref_mha(xq64, xkv64, w, h, 20, 500, 512, 0.1, seeds)
    (want * gy[0].double()).sum().backward()
    assert G.rel_err(y[0].cpu(), want.detach()) < 1e-3
    assert G.rel_err(q_dev.grad[0, :, :, 0].t().cpu(), xq64.grad) < 1e-3
    assert G.rel_err(kv_dev.grad[0, :, :, 0].t().cpu(), xkv64.grad) < 1e-3
    for name, prm in m.named_parameters():
        assert G.rel_err(prm.grad.cpu(), w[name].grad) < 1e-3, name
    # eval mode is untouched by the seed; train mode differs from it and is reproducible under torch.manual_seed
    m.eval()
    with torch.no_grad():
        e1, _ = m(q_dev, kv_dev, kv_dev, "test")
        e2, _ = m(q_dev, kv_dev, kv_dev, "test")
    assert torch.equal(e1, e2)
    m.train()
    with torch.no_grad():
        torch.manual_seed(5); t1, _ = m(q_dev, kv_dev, kv_dev, "train")
        torch.manual_seed(5); t2, _ = m(q_dev, kv_dev, kv_dev, "train")
        torch.manual_seed(6); t3, _ = m(q_dev, kv_dev, kv_dev, "train")
    assert torch.equal(t1, t2) and not torch.equal(t1, t3)
    assert G.rel_err(t1, e1) > 1e-2


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-3), ("bf16", 1e-2)])
def test_mink_mha_train_mode_matches_reference_with_the_same_masks(precision, tol):
    """d_head = 64, h = 4, full (unchunked) attention with Lq != Lk: the 128-key kernels."""
    from csn_b200 import mink
    h, seed, Lq, Lk = 4, 31, 300, 200
    sd = {k[len("MHA."):]: v for k, v in synth.mink_state(seed, h).items() if k.startswith("MHA.")}
    m = mink.MultiHeadAttention(h, 256, 64, 64, precision=precision).cuda().train()
    m.load_state_dict(sd)
    gen = synth.gen(seed + 1)
    q = torch.relu(torch.randn(1, Lq, 256, generator=gen))
    k = torch.relu(torch.randn(1, Lk, 256, generator=gen))
    gy = torch.randn(1, Lq, 256, generator=gen)
    seeds = _seeds(77)
    torch.manual_seed(77)
    qd, kd = q.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    out, _ = m(qd, kd, kd)
    (out * gy.cuda()).sum().backward()
    w = {n: v.double().requires_grad_(True) for n, v in sd.items()}
    q64, k64 = q[0].double().requires_grad_(True), k[0].double().requires_grad_(True)
    n_pad = (max(Lq, Lk) + 127) // 128 * 128
    want = _ref_mha(q64, k64, w, h, 1, max(Lq, Lk), n_pad, 0.1, seeds)[:Lq] if Lq == Lk else None
    if want is None:   # Lq != Lk: one chunk holding all rows of both sides
        d = 64
        Q = (q64 @ w["w_qs.weight"].t()).view(Lq, h, d).transpose(0, 1)
        K = (k64 @ w["w_ks.weight"].t()).view(Lk, h, d).transpose(0, 1)
        V = (k64 @ w["w_vs.weight"].t()).view(Lk, h, d).transpose(0, 1)
        P = torch.softmax((Q / d ** 0.5) @ K.transpose(1, 2), dim=-1)
        masks = torch.stack([keep_mask(seeds[0], a * n_pad + np.arange(Lq), Lk, 0.1) for a in range(h)])
        O = ((P * masks) @ V).transpose(0, 1).reshape(Lq, h * d)
        fc = (O @ w["fc.weight"].t()) * keep_mask(seeds[1], np.arange(Lq), 256, 0.1)
        want = F.layer_norm(fc + q64, (256,), w["norm.weight"], w["norm.bias"], 1e-6)
    (want * gy[0].double()).sum().backward()
    assert G.rel_err(out[0].cpu(), want.detach()) < tol
    assert G.rel_err(qd.grad[0].cpu(), q64.grad) < tol
    assert G.rel_err(kd.grad[0].cpu(), k64.grad) < tol
    for name, prm in m.named_parameters():
        assert G.rel_err(prm.grad.cpu(), w[name].grad) < tol, name


def test_mask_statistics():
    """Drop rate of the hash and mean preservation of the scaled mask (nn.Dropout's two defining properties)."""
    msk = keep_mask(12345, np.arange(4000), 512, 0.1)
    rate = float((msk == 0).double().mean())
    assert abs(rate - 0.1) < 2e-3, rate                       # 2M samples: sigma = 2e-4
    assert abs(float(msk.mean()) - 1.0) < 2e-3
    # rows and columns are uncorrelated: per-row / per-column drop rates stay within 5 sigma of p
    assert float(((msk == 0).double().mean(1) - 0.1).abs().max()) < 5 * (0.09 / 512) ** 0.5
    assert float(((msk == 0).double().mean(0) - 0.1).abs().max()) < 5 * (0.09 / 4000) ** 0.5
    # different seeds give independent masks
    other = keep_mask(12346, np.arange(4000), 512, 0.1)
    both = float(((msk == 0) & (other == 0)).double().mean())
    assert abs(both - 0.01) < 1e-3


def test_csa_training_step_in_train_mode():
    """CrossShapeAt in model.train() (what csa_training.py:191 does): the fused step runs with dropout on, is
    reproducible under torch.manual_seed, differs from eval, and its loss stays close to the eval loss."""
    from csn_b200 import midfc
    B, K, h, C = 2, 2, 1, 15
    m = midfc.get_model("csa", C, h, K).cuda()
    m.load_state_dict(synth.midfc_state(3, h, C))
    x, nb = synth.csa_batch(4, B, K)
    lab = torch.randint(0, C, (B, x.shape[2]), generator=synth.gen(5)).cuda()
    x, nb = x.cuda(), nb.cuda()
    m.eval()
    le = m.forward_loss(x, "test", nb, lab)
    le.backward()
    ge = m.attention.w_qs.weight.grad.clone()
    m.zero_grad()
    m.train()
    res = []
    for s in (1, 1, 2):
        torch.manual_seed(s)
        lt = m.forward_loss(x, "train", nb, lab)
        lt.backward()
        res.append((lt.item(), m.attention.w_qs.weight.grad.clone()))
        m.zero_grad()
    assert abs(res[0][0] - res[1][0]) < 1e-6 and G.rel_err(res[0][1], res[1][1]) < 1e-5   # (split-K sums: order varies)
    assert res[0][0] != res[2][0]
    assert abs(res[0][0] - le.item()) < 0.05 * abs(le.item()) and res[0][0] != le.item()
    assert torch.isfinite(res[0][1]).all() and G.rel_err(res[0][1], ge) < 1.0
    # the unfused module path (model(x) -> logits) honours train mode as well
    torch.manual_seed(1)
    logits_t = m(x, "train", nb)
    m.eval()
    logits_e = m(x, "test", nb)
    assert G.rel_err(logits_t, logits_e) > 1e-3 and torch.isfinite(logits_t).all()


def test_graph_replay_draws_fresh_masks_per_epoch():
    """A train-mode step captured as a CUDA graph freezes its seeds; the device-resident epoch word offsets them:
    same epoch -> bit-identical loss, different epochs -> different masks, epoch 0 -> the eager result."""
    from csn_b200 import midfc, graphs
    torch.manual_seed(3)
    m = midfc.get_model("csa", 15, 1, 2).cuda().train()
    m.load_state_dict(synth.midfc_state(5, 1, 15))
    x, nb = synth.csa_batch(6, 2, 2)
    x, nb = x.cuda(), nb.cuda()
    lab = torch.randint(1, 15, (2, x.shape[2]), generator=torch.Generator().manual_seed(1)).cuda()
    params = [p_ for n, p_ in m.named_parameters() if not n.startswith("fc_1")]

    def step(x, nb, lab):
        for p_ in params:
            p_.grad = None
        loss = m.forward_loss(x, "train", nb, lab)
        loss.backward()
        return loss

    g = graphs.GraphedStep(step, x, nb, lab)
    try:
        l0 = g.replay(epoch=0).item()
        l0b = g.replay(epoch=0).item()
        l1 = g.replay(epoch=1).item()
        l2 = g.replay(epoch=2).item()
        g1 = m.attention.w_qs.weight.grad.clone()
        l1b = g.replay(epoch=1).item()
        assert l0 == l0b and l1 == l1b
        assert len({l0, l1, l2}) == 3
        assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
        # all within dropout noise of each other
        assert max(abs(l1 - l0), abs(l2 - l0)) < 0.2 * abs(l0)
    finally:
        graphs.set_drop_epoch(0)
