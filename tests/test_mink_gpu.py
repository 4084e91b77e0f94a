"""GPU parity of the MinkowskiNet attention surface (csn_b200.mink) against the reference golden
vectors (MinkowskiNet/models/attention.py run by oracle/make_golden.py) and the CPU oracle."""
import pytest
import torch

from csn_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _mha(seed, h, precision="fp16", return_attn=False):
    from csn_b200 import mink
    m = mink.MultiHeadAttention(h, 256, 256 // h, 256 // h, precision=precision, return_attn=return_attn).cuda().eval()
    sd = synth.mink_state(seed, h)
    m.load_state_dict({k[len("MHA."):]: v for k, v in sd.items() if k.startswith("MHA.")})
    return m


def test_mha_forward_backward_matches_reference():
    g = G.load("mink_mha")
    seed, h, Lq, Lk = (int(g[k]) for k in ("seed", "n_head", "Lq", "Lk"))
    m = _mha(seed, h, return_attn=True)
    gen = synth.gen(seed + 1)
    q = torch.relu(torch.randn(1, Lq, 256, generator=gen)).cuda().requires_grad_(True)
    k = torch.relu(torch.randn(1, Lk, 256, generator=gen)).cuda().requires_grad_(True)
    out, attn = m(q, k, k)
    gy = torch.randn(out.shape, generator=gen).cuda()
    (out * gy).sum().backward()
    assert out.shape == (1, Lq, 256) and attn.shape == (1, h, Lq, Lk)
    G.compare_sampled(g, "out", out, TOL)
    G.compare_sampled(g, "attn", attn, 4 * TOL)
    G.compare_sampled(g, "grad.q", q.grad, TOL)
    G.compare_sampled(g, "grad.k", k.grad, TOL)
    for pname in ("w_qs.weight", "w_ks.weight", "w_vs.weight", "fc.weight", "norm.weight", "norm.bias"):
        G.compare_sampled(g, "grad." + pname, dict(m.named_parameters())[pname].grad, TOL, what=pname)


def test_ragged_self_attention_against_oracle():
    """Lengths that are not multiples of 128, q is k (SSA call pattern hrnet.py:461-463)."""
    from oracle import csa_oracle as O
    h = 4
    m = _mha(7, h)
    w = synth.mink_state(7, h)
    for L in (1, 5, 37, 64, 65, 129, 1000):
        x = torch.relu(torch.randn(1, L, 256, generator=synth.gen(L)))
        want, _ = O.mha_mink(x, x, x, w, h)
        got, attn = m(x.cuda(), x.cuda(), x.cuda())
        assert attn is None
        assert G.rel_err(got.cpu(), want) < TOL, L


def test_csa_head_against_oracle():
    """CSA block of HRNetSimCSN.forward (hrnet.py:370-417) on a ragged batch, K = 2."""
    from csn_b200 import mink
    from oracle import csa_oracle as O
    h = 4
    head = mink.CSAHead(256, h).cuda().eval()
    sd = synth.mink_state(9, h)
    head.load_state_dict(sd)
    gen = synth.gen(10)
    lens_q, lens_k = [150, 333], [[200, 129], [90, 400]]
    qf = [torch.relu(torch.randn(n, 256, generator=gen)) for n in lens_q]
    kf = [[torch.relu(torch.randn(n, 256, generator=gen)) for n in ln] for ln in lens_k]
    want = O.mink_csa_block(qf, kf, sd, h)
    with torch.no_grad():
        got = head([t.cuda() for t in qf], [[t.cuda() for t in lst] for lst in kf])
    for a, b in zip(got, want):
        assert G.rel_err(a.cpu(), b) < TOL
    sim = mink.cosine_similarity(qf[0].cuda(), kf[0][1].cuda()).item()
    assert abs(sim - O.mink_cosine_similarity(qf[0], kf[0][1]).item()) < 1e-5
    s = torch.tensor([0.1, 0.9, 0.5, 0.7]).cuda()
    assert mink.topk_neighbors(s, 2, self_index=1).tolist() == O.mink_topk_neighbors(s.cpu(), 2, self_index=1).tolist()


def test_batched_csa_head_matches_call_by_call_loop():
    """The ragged one-pass batch (MultiHeadAttention.forward_blocks: per-block lengths in the kernels' work
    tables) against the reference's call-by-call loop: outputs, input gradients, weight gradients."""
    from csn_b200 import mink
    h = 4
    head = mink.CSAHead(256, h).cuda().eval()
    head.load_state_dict(synth.mink_state(11, h))
    gen = synth.gen(12)
    lens_q, lens_k = [300, 77, 512], [[129, 260, 40], [333, 128, 700]]
    res = {}
    for batched in (True, False):
        g2 = synth.gen(12)
        qf = [torch.relu(torch.randn(n, 256, generator=g2)).cuda().requires_grad_(True) for n in lens_q]
        kf = [[torch.relu(torch.randn(n, 256, generator=g2)).cuda().requires_grad_(True) for n in ln] for ln in lens_k]
        head.zero_grad()
        out = head(qf, kf, batched=batched)
        sum((o * torch.linspace(-1, 1, 256, device="cuda")).sum() for o in out).backward()
        res[batched] = ([o.detach() for o in out], [t.grad.clone() for t in qf] + [t.grad.clone() for lst in kf for t in lst],
                        head.MHA.w_qs.weight.grad.clone(), head.MHA.fc.weight.grad.clone(), head.linear_k.weight.grad.clone())
    for a, b in zip(res[True][0], res[False][0]):
        assert G.rel_err(a, b) < 2e-4
    for a, b in zip(res[True][1], res[False][1]):
        assert G.rel_err(a, b) < 1e-3
    for a, b in zip(res[True][2:], res[False][2:]):
        assert G.rel_err(a, b) < 1e-3


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-3), ("bf16", 1e-2)])
def test_csa_head_forward_backward_matches_reference(precision, tol):
    """CSA block of HRNetSimCSN.forward (hrnet.py:370-417), forward AND backward, against vectors produced with the
    reference's own MultiHeadAttention module (tests/golden/mink_csa_head.npz, oracle/make_golden.py): features,
    gradients w.r.t. every input shape (the backbone trains end to end) and every parameter.  bf16 = config 4's dtype."""
    from csn_b200 import mink
    g = G.load("mink_csa_head")
    seed, h, K = int(g["seed"]), int(g["n_head"]), int(g["K"])
    lens, key_lens = [int(v) for v in g["lens"]], [[int(v) for v in row] for row in g["key_lens"]]
    head = mink.CSAHead(256, h, precision=precision).cuda().eval()
    head.load_state_dict(synth.mink_state(seed, h))
    gen = synth.gen(seed + 1)
    qf = [torch.relu(torch.randn(L, 256, generator=gen)).cuda().requires_grad_(True) for L in lens]
    kf = [[torch.relu(torch.randn(L, 256, generator=gen)).cuda().requires_grad_(True) for L in kl] for kl in key_lens]
    gys = [torch.randn(L, 256, generator=gen).cuda() for L in lens]
    outs = head(qf, kf)
    sum((o * gy).sum() for o, gy in zip(outs, gys)).backward()
    for b, o in enumerate(outs):
        G.compare_sampled(g, f"out{b}", o, tol)
        G.compare_sampled(g, f"grad.q{b}", qf[b].grad, tol)
        for i in range(K):
            G.compare_sampled(g, f"grad.k{i}_{b}", kf[i][b].grad, tol)
    params = dict(head.named_parameters())
    for pname in ("MHA.w_qs.weight", "MHA.w_ks.weight", "MHA.w_vs.weight", "MHA.fc.weight", "MHA.norm.weight",
                  "MHA.norm.bias"):
        G.compare_sampled(g, "grad." + pname, params[pname].grad, tol, what=pname)
    # linear_q / linear_k see the same cancellation as MID-FC's compatibility layers: norm-scaled bound
    scale = max(float(g[f"grad.MHA.{n}.sumsq"]) ** 0.5 for n in ("w_qs.weight", "fc.weight"))
    for pname in ("linear_q.weight", "linear_k.weight"):
        stride = int(g[f"grad.{pname}.stride"])
        got = params[pname].grad.detach().reshape(-1).double().cpu()[::stride]
        want = torch.from_numpy(g[f"grad.{pname}.values"].astype("float64"))
        assert float((got - want).norm()) < tol * scale, pname


def test_gradients_follow_autograd_identity_not_storage():
    """ADVICE (round 1): de-duplication of q / k / v slots is decided on object identity.  (a) distinct k and v tensors
    holding the same data each receive their own gradient (their sum is the gradient of the shared-tensor call);
    (b) k = q.detach() shares q's storage but is a different autograd leaf: q receives only the query-role gradient."""
    m = _mha(7, 4)
    gen = synth.gen(3)
    q = torch.relu(torch.randn(1, 200, 256, generator=gen)).cuda()
    kv = torch.relu(torch.randn(1, 150, 256, generator=gen)).cuda()
    gy = torch.randn(1, 200, 256, generator=gen).cuda()
    # shared tensor: reference gradient of the key/value input
    k0 = kv.clone().requires_grad_(True)
    (m(q, k0, k0)[0] * gy).sum().backward()
    # (a) distinct tensors
    k1, v1 = kv.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    (m(q, k1, v1)[0] * gy).sum().backward()
    assert v1.grad is not None and float(v1.grad.abs().max()) > 0 and float(k1.grad.abs().max()) > 0
    assert G.rel_err(k1.grad + v1.grad, k0.grad) < 2e-4
    assert G.rel_err(k1.grad, k0.grad) > 1e-2          # the value-role part is no longer credited to k
    # (b) alias of q's storage
    q2 = q.clone().requires_grad_(True)
    (m(q2, q2.detach(), q2.detach())[0] * gy).sum().backward()
    q3, k3 = q.clone().requires_grad_(True), q.clone()
    (m(q3, k3, k3)[0] * gy).sum().backward()
    assert G.rel_err(q2.grad, q3.grad) < 2e-4
    q4 = q.clone().requires_grad_(True)
    (m(q4, q4, q4)[0] * gy).sum().backward()          # true self-attention: all three roles
    assert G.rel_err(q4.grad, q3.grad) > 1e-2


def test_sparse_glue_split_by_batch_and_head():
    """features_at for all items at once (lib/utils.py:283-288) and the CSA block fed with (.F, .C[:, 0]) pairs: grouped
    rows give views, shuffled rows are gathered; the stacked output equals the per-item path."""
    from csn_b200 import mink
    gen = synth.gen(21)
    lens = [150, 90, 260]
    feats = torch.relu(torch.randn(sum(lens), 256, generator=gen)).cuda()
    bcol = torch.cat([torch.full((n,), b) for b, n in enumerate(lens)]).cuda()
    parts, got_lens, perm = mink.split_by_batch(feats, bcol)
    assert got_lens == lens and perm is None
    assert parts[1].data_ptr() == feats[150:].data_ptr()                      # a view, not a copy
    for b in range(3):
        assert torch.equal(parts[b], feats[bcol == b])                        # == features_at(sparse, b)
    shuffle = torch.randperm(sum(lens), generator=gen).cuda()
    parts2, _, perm2 = mink.split_by_batch(feats[shuffle], bcol[shuffle])
    for b in range(3):
        assert torch.equal(parts2[b], feats[shuffle][bcol[shuffle] == b])
    parts3, _, _ = mink.split_by_batch(feats, bcol, lens=lens)               # lens from the loader: no device sync
    assert all(torch.equal(a, b) for a, b in zip(parts, parts3))
    head = mink.CSAHead(256, 4).cuda().eval()
    head.load_state_dict(synth.mink_state(9, 4))
    klens = [[120, 200, 64]]
    kfeats = torch.relu(torch.randn(sum(klens[0]), 256, generator=gen)).cuda()
    kcol = torch.cat([torch.full((n,), b) for b, n in enumerate(klens[0])]).cuda()
    with torch.no_grad():
        got = mink.csa_head_sparse(head, feats, bcol, [(kfeats, kcol)])
        want = torch.cat(head(list(torch.split(feats, lens)), [list(torch.split(kfeats, klens[0]))]), dim=0)
    assert got.shape == (sum(lens), 256) and G.rel_err(got, want) < 1e-5


def test_construct_shape_graph_against_oracle():
    """lib/csn_utils.py:46-100 with a GPU-resident key store: neighbour lists equal the oracle's pair-by-pair loop
    (mink_cosine_similarity + mink_topk_neighbors), with and without self-exclusion."""
    from csn_b200 import mink
    from oracle import csa_oracle as O
    h, K = 4, 2
    head = mink.CSAHead(256, h).cuda().eval()
    sd = synth.mink_state(13, h)
    head.load_state_dict(sd)
    gen = synth.gen(14)
    protos = torch.randn(3, 256, generator=gen)
    shapes = [torch.relu(protos[i % 3] + 0.7 * torch.randn(int(n), 256, generator=gen)) for i, n in enumerate([140, 90, 200, 77, 160, 128, 111])]
    ssa = O.mink_ssa(shapes, sd, h)
    sim = torch.stack([torch.stack([O.mink_cosine_similarity(a, b) for b in ssa]) for a in ssa])
    dev_shapes = [s.cuda() for s in shapes]
    got = mink.construct_shape_graph(head, dev_shapes, dev_shapes, K, is_same=True)
    for q, nb in got:
        assert nb == O.mink_topk_neighbors(sim[q], K, self_index=q).tolist(), (q, nb)
    got2 = mink.construct_shape_graph(head, dev_shapes[:3], dev_shapes[3:], K, is_same=False)
    for q, nb in got2:
        assert nb == sim[q, 3:].topk(K).indices.tolist(), (q, nb)


@pytest.mark.parametrize("precision,tol", [("fp16", 1.5e-3), ("bf16", 1.2e-2)])
def test_fused_dkv_backward_matches_ds_path(precision, tol):
    """d_head 64: the key-stationary dK/dV kernel (csn_attn_bwd_dkv; dS never materialised) against the older
    dV kernel + dS buffer + dK GEMM path on a ragged batch with cross blocks (both are separately pinned to the
    reference by the golden tests above; here they must agree with each other block by block)."""
    import os
    from csn_b200 import mink
    h = 4
    m = mink.MultiHeadAttention(h, 256, 64, 64, precision=precision).cuda().eval()
    m.load_state_dict({k[len("MHA."):]: v for k, v in synth.mink_state(11, h).items() if k.startswith("MHA.")})
    gen = synth.gen(12)
    lens = [130, 517, 64, 300, 1, 33]          # incl. shapes shorter than one column quarter of a tile
    pairs = [(0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (5, 5), (0, 1), (1, 3), (2, 0), (4, 1), (1, 5), (5, 4)]
    xs = [torch.relu(torch.randn(n, 256, generator=gen)).cuda() for n in lens]
    gy = [torch.randn(lens[q], 256, generator=gen).cuda() for q, _ in pairs]

    def run(flag):
        os.environ["CSN_FUSED_DKV"] = flag
        try:
            leaves = [x.clone().requires_grad_(True) for x in xs]
            for p_ in m.parameters():
                p_.grad = None
            ys = m.forward_blocks(leaves, pairs)
            sum((y * g).sum() for y, g in zip(ys, gy)).backward()
            return [t.grad.clone() for t in leaves] + [m.w_qs.weight.grad.clone(), m.w_ks.weight.grad.clone(),
                                                       m.w_vs.weight.grad.clone()]
        finally:
            os.environ.pop("CSN_FUSED_DKV", None)

    a, b = run("1"), run("0")
    for i, (x, y) in enumerate(zip(a, b)):
        assert torch.isfinite(x).all()
        assert G.rel_err(x, y) < tol, (i, G.rel_err(x, y))


@pytest.mark.parametrize("L", [300, 1000])
def test_growing_scores_force_accumulator_rescale(L):
    """Keys whose scores grow along the sequence: every later key tile raises the running row maximum by far more
    than the 2^8 the kernels tolerate, so the lazy re-scale of the TMEM accumulators (rare on random data) runs in
    (almost) every tile, also in the second column half.  Forward + backward against the oracle."""
    from oracle import csa_oracle as O
    h = 4
    m = _mha(21, h)
    w = synth.mink_state(21, h)
    g = synth.gen(22)
    x = torch.relu(torch.randn(1, L, 256, generator=g))
    x = x * torch.linspace(0.2, 6.0, L).view(1, L, 1)          # row norms (and with them the scores) grow with the index
    xr = x.clone().requires_grad_(True)
    want, _ = O.mha_mink(xr, xr, xr, w, h)
    gy = torch.randn(want.shape, generator=g)
    (want * gy).sum().backward()
    xd = x.cuda().requires_grad_(True)
    got, _ = m(xd, xd, xd)
    (got * gy.cuda()).sum().backward()
    assert torch.isfinite(got).all()
    assert G.rel_err(got.detach().cpu(), want.detach()) < TOL
    assert G.rel_err(xd.grad.cpu(), xr.grad) < 2e-3
