"""CPU oracle for the CSN cross-shape-attention hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A from-scratch, functional restatement (PyTorch CPU, fp32 or fp64) of the reference arithmetic that
the CUDA kernels in csn_b200/csrc must reproduce.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product path
(csn_b200/*) never does and fails loudly without its CUDA library.

Pinning: the reference has NO tests, golden vectors or fixtures for this path,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports
/root/reference/MID-FC/csa_models.py and /root/reference/MinkowskiNet/models/attention.py in the
build container, runs them on the seeded inputs of csn_b200/synth.py and commits sub-sampled
outputs under tests/golden/; tests/test_oracle_golden.py checks this file against those vectors
(and, when /root/reference is present, against the live reference).

Every function cites the reference lines it restates (paths relative to /root/reference).
Weights are passed as a plain dict using the reference's state_dict key names.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # MID-FC/csa_models.py:57, MinkowskiNet/models/attention.py:29


# --------------------------------------------------------------------------- attention core
def scaled_dot_product_attention(q, k, v, temperature):
    """MID-FC/csa_models.py:138-144 == MinkowskiNet/models/attention.py:69-75 (dropout off).
    q,k,v: (..., L, d).  The division is applied to q BEFORE the product (csa_models.py:139)."""
    scores = torch.matmul(q / temperature, k.transpose(-1, -2))
    attn = torch.softmax(scores, dim=-1)
    return torch.matmul(attn, v), attn


def _project_heads(x, w, n_head):
    """nn.Linear(bias=False) then split heads: (B, L, D) -> (B, h, L, d)  (csa_models.py:103-108)."""
    B, L, _ = x.shape
    y = x @ w.t()
    return y.view(B, L, n_head, -1).transpose(1, 2)


def _mha_rows(xq, xkv, w, n_head, prefix):
    """One un-chunked MHA over row-major inputs (B, Lq, D), (B, Lk, D):
    LayerNorm(fc(concat_heads(softmax(QK^T/sqrt(d_k)) V)) + xq).
    MinkowskiNet/models/attention.py:31-56; also the body of one chunk of csa_models.py:96-118."""
    wq, wk, wv = w[prefix + "w_qs.weight"], w[prefix + "w_ks.weight"], w[prefix + "w_vs.weight"]
    wo = w[prefix + "fc.weight"]
    d_k = wq.shape[0] // n_head
    q = _project_heads(xq, wq, n_head)
    k = _project_heads(xkv, wk, n_head)
    v = _project_heads(xkv, wv, n_head)
    o, attn = scaled_dot_product_attention(q, k, v, math.sqrt(d_k))
    B, _, Lq, _ = o.shape
    o = o.transpose(1, 2).reshape(B, Lq, -1)
    z = o @ wo.t() + xq
    y = F.layer_norm(z, (z.shape[-1],), w[prefix + "norm.weight"], w[prefix + "norm.bias"], LN_EPS)
    return y, attn


def mha_midfc(Q, K, V, w, n_head, prefix="attention.", iters=20, chunk=500):
    """MID-FC/csa_models.py:81-125.  Q,K,V channel-major (B, D, N, 1).  Block-diagonal attention:
    rows [500c, 500c+500) of Q attend only to the same rows of K/V, 20 chunks (N must be >= 10000;
    points beyond 10000 are silently dropped, fewer raise IndexError — SURVEY F6).  The reference
    passes K and V separately but every caller passes the same tensor.  Returns
    (Y (B, iters*chunk, D) row-major, attention matrix of the LAST chunk (B, h, chunk, chunk))."""
    n_used = iters * chunk
    if Q.shape[2] < n_used or K.shape[2] < n_used or V.shape[2] < n_used:
        raise IndexError(f"MID-FC attention indexes points [0,{n_used}); got N={Q.shape[2]}")
    if K.data_ptr() != V.data_ptr() and not torch.equal(K, V):
        raise NotImplementedError("oracle restates the K is V call pattern of every reference caller")
    B, D = Q.shape[0], Q.shape[1]
    xq = Q[:, :, :n_used, 0].transpose(1, 2)  # (B, N, D) row-major view
    xkv = K[:, :, :n_used, 0].transpose(1, 2)
    # fold chunks into the batch dimension: (B*iters, chunk, D)
    xq_c = xq.reshape(B * iters, chunk, D)
    xkv_c = xkv.reshape(B * iters, chunk, D)
    y, attn = _mha_rows(xq_c, xkv_c, w, n_head, prefix)
    y = y.reshape(B, n_used, D)
    attn_last = attn.reshape(B, iters, n_head, chunk, chunk)[:, -1]
    return y, attn_last


def mha_mink(q, k, v, w, n_head, prefix="MHA."):
    """MinkowskiNet/models/attention.py:31-56: full Lq x Lk attention, row-major (B, L, D)."""
    if k.data_ptr() != v.data_ptr() and not torch.equal(k, v):
        raise NotImplementedError("oracle restates the k is v call pattern (hrnet.py:407,463)")
    return _mha_rows(q, k, w, n_head, prefix)


# --------------------------------------------------------------------------- MID-FC CSA layer
def _to_channel_major(y):
    """(B, N, D) -> (B, D, N, 1)   (csa_models.py:206,240)."""
    return y.transpose(1, 2).unsqueeze(-1)


def ssa_feats(x, w, n_head):
    """CrossShapeAt.get_ssa_feats, csa_models.py:204-207."""
    y, attn = mha_midfc(x, x, x, w, n_head)
    return _to_channel_major(y), attn


def compatibility(y_q, y_stack, w, batch):
    """csa_models.py:222-230 including the batch-interleaving view:
    y_stack rows are ordered [k=0: b=0..B-1; k=1: b=0..B-1; ...] and then VIEWED as (B, K+1, D)."""
    u_q = F.normalize(y_q @ w["compatibility_q.weight"].t() + w["compatibility_q.bias"], dim=-1)
    u_k = F.normalize(y_stack @ w["compatibility_k.weight"].t() + w["compatibility_k.bias"], dim=-1)
    u_k = u_k.view(batch, -1, u_k.shape[1])
    logits = torch.einsum("bd,bkd->bk", u_q, u_k)
    return torch.softmax(logits, dim=-1)


def csa_feats(x, x_neighbors, w, n_head):
    """CrossShapeAt.get_csa_feats, csa_models.py:209-242.  x (B,D,N,1); x_neighbors
    (B,K+1,D,N,1), slot 0 is the query itself and is skipped (:214,:234)."""
    B = x.shape[0]
    y_self, _ = mha_midfc(x, x, x, w, n_head)  # SSA(x); reused at :232 (identical in eval mode)
    pooled = [y_self.mean(dim=1)]
    for k in range(1, x_neighbors.shape[1]):
        xk = x_neighbors[:, k]
        yk, _ = mha_midfc(xk, xk, xk, w, n_head)
        pooled.append(yk.mean(dim=1))
    comp = compatibility(pooled[0], torch.cat(pooled, dim=0), w, B)
    out = comp[:, 0, None, None] * y_self
    for k in range(1, x_neighbors.shape[1]):
        xk = x_neighbors[:, k]
        cross, _ = mha_midfc(x, xk, xk, w, n_head)
        out = out + comp[:, k, None, None] * cross
    return _to_channel_major(out)


def logits_1x1(feats, w):
    """self.logit: 1x1 conv 256 -> num_classes, no bias (csa_models.py:151,178,201)."""
    return F.conv2d(feats, w["logit.weight"])


def forward_csa(x, x_neighbors, w, n_head):
    """CrossShapeAt.forward with attention_type='csa', after_fc=True (csa_models.py:182-202)."""
    return logits_1x1(csa_feats(x, x_neighbors, w, n_head), w)


def forward_ssa(x, w, n_head):
    """attention_type='ssa', after_fc=True (csa_models.py:191-195)."""
    return logits_1x1(ssa_feats(x, w, n_head)[0], w)


def masked_cross_entropy(logits, label):
    """Loss of the training step used by config 2 (MID-FC/csa_training.py:94-108): points with
    label 0 are ignored, mean CE over the rest.  logits (B,C,N,1), label (B,N) int64."""
    C = logits.shape[1]
    lg = logits.squeeze(-1).permute(0, 2, 1).reshape(-1, C)
    lb = label.reshape(-1)
    keep = lb > 0
    return F.cross_entropy(lg[keep], lb[keep])


# --------------------------------------------------------------------------- retrieval / kNN graph
def retrieval_measure(f1, f2, eps=1e-12, block=2048):
    """CrossShapeAt.get_retrieval_measure, csa_models.py:244-267 (big-class twin :360-392):
    score[i,j] = mean_p max_q cos(f1[i,p], f2[j,q]); rows L2-normalised with F.normalize
    (v / max(|v|, 1e-12)).  f1 (Sq,N,D), f2 (Sc,M,D) -> (Sq,Sc).  The (N,M) cosine matrix is
    processed in row blocks so the oracle also runs at N=10000 without 400 MB temporaries."""
    a = F.normalize(f1, dim=-1, eps=eps)
    b = F.normalize(f2, dim=-1, eps=eps)
    Sq, Sc = a.shape[0], b.shape[0]
    out = torch.empty(Sq, Sc, dtype=f1.dtype)
    for i in range(Sq):
        for j in range(Sc):
            acc = a.new_zeros(())
            for r in range(0, a.shape[1], block):
                acc = acc + (a[i, r:r + block] @ b[j].t()).max(dim=-1).values.sum()
            out[i, j] = acc / a.shape[1]
    return out


def knn_graph(f1, f2, K):
    """CrossShapeAt.get_knn_graph, csa_models.py:270-280: topk(K+1) indices, sorted descending,
    self included when query set == candidate set."""
    return retrieval_measure(f1, f2).topk(K + 1, dim=-1).indices


# --------------------------------------------------------------------------- MinkowskiNet CSA head
def mink_cosine_similarity(q, k):
    """HRNetSimCSN.cosine_similarity, MinkowskiNet/models/hrnet.py:472-490: rows divided by their
    raw L2 norm (no eps), mean over q rows of the max over k rows."""
    qn = q / q.pow(2).sum(dim=1).sqrt().unsqueeze(1)
    kn = k / k.pow(2).sum(dim=1).sqrt().unsqueeze(1)
    return (qn @ kn.t()).max(dim=1).values.mean()


def mink_ssa(feats, w, n_head):
    """HRNetSimCSN.get_SSA (hrnet.py:456-470) over a list of per-shape (L_b, D) tensors (the dense
    stand-in for features_at on an ME.SparseTensor, lib/utils.py:283-288)."""
    return [mha_mink(f[None], f[None], f[None], w, n_head)[0][0] for f in feats]


def mink_csa_block(query_feats, key_feats, w, n_head):
    """CSA block of HRNetSimCSN.forward, hrnet.py:370-417.  query_feats: list over batch items of
    (L_b, D); key_feats: list over K neighbours of such lists.  Returns the list of per-item CSA
    features.  linear_q/linear_k have no bias (:355-356); similarity temperature sqrt(d_model)
    (:357)."""
    d_model = query_feats[0].shape[1]
    q_ssa = mink_ssa(query_feats, w, n_head)
    keys_ssa = [q_ssa] + [mink_ssa(kf, w, n_head) for kf in key_feats]
    out = []
    for b, ssa_b in enumerate(q_ssa):
        g_q = F.normalize(ssa_b.mean(dim=0) @ w["linear_q.weight"].t(), dim=-1)
        sims = []
        for ks in keys_ssa:
            g_k = F.normalize(ks[b].mean(dim=0) @ w["linear_k.weight"].t(), dim=-1)
            sims.append((g_q * g_k).sum() / math.sqrt(d_model))
        comp = torch.softmax(torch.stack(sims), dim=0)
        csa = comp[0] * ssa_b
        for i, kf in enumerate(key_feats):
            cross, _ = mha_mink(query_feats[b][None], kf[b][None], kf[b][None], w, n_head)
            csa = csa + comp[i + 1] * cross[0]
        out.append(csa)
    return out


def mink_topk_neighbors(sim_row, K, self_index=None):
    """MinkowskiNet/lib/csn_utils.py:91-96: topk(K); if the query itself is among them, redo with
    topk(K+1) and drop it."""
    idx = sim_row.topk(K).indices
    if self_index is not None and bool((idx == self_index).any()):
        idx = sim_row.topk(K + 1).indices
        idx = idx[idx != self_index]
    return idx
