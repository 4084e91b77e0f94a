"""Drop-in module surface of MID-FC/csa_models.py backed by the sm_100a kernels.

Same class names, constructor signatures, state_dict keys, argument meaning and return shapes as the
reference (MID-FC/csa_models.py:37-432; contract in SURVEY.md §8b):
    MultiHeadAttention, ScaledDotProductAttention, CrossShapeAt, get_model
Tensors in/out are fp32 in the reference layouts (channel-major (B,256,N,1) features); the
contractions run on tcgen05 tensor cores with 16-bit operands and fp32 accumulation
(`precision='fp16'`: 11-bit mantissa like TF32, default; `'bf16'`: the wide-range variant).
There is no fallback: without a CUDA device or without libcsn_b200.so every call raises.

Deviations from the reference, all documented in DESIGN.md:
  * dropout (csa_models.py:56,136) is not applied: results are those of `model.eval()`;
  * projections of a shape are computed once per step instead of once per attention call, and the
    duplicate self-attention call of get_csa_feats (:210 vs :232) is computed once.
"""
from __future__ import annotations

import numpy as np
import contextlib
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import engine as E
from . import knn as _knn

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")  # csa_models.py:8 (module-level)

_PRECISIONS = {"fp16": torch.float16, "bf16": torch.bfloat16}


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise L.CsnError(f"{what} must be a CUDA tensor: csn_b200 has no CPU path")


def neighbors_to_device(x_neighbors: torch.Tensor, dev, stream: "torch.cuda.Stream" = None) -> torch.Tensor:
    """Host -> device copy of a neighbour tensor (B,K+1,256,N,1) that skips slot 0: the layer never reads it
    (csa_models.py:214,234 iterate k = 1..K), so a quarter of the bytes (K=3) stay off the PCIe link. The device
    tensor keeps the (B,K+1,...) shape with slot 0 left uninitialised. With pinned host memory the copies are
    asynchronous; `stream` (optional) lets an input pipeline stage step i+1 while step i computes."""
    out = torch.empty(x_neighbors.shape, dtype=x_neighbors.dtype, device=dev)
    with torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext():
        for b in range(x_neighbors.shape[0]):
            out[b, 1:].copy_(x_neighbors[b, 1:], non_blocking=True)
    return out


def _geom_for(n_src_points: int, iters: int = 20, chunk: int = 500) -> E.Geometry:
    g = E.Geometry(chunk=chunk, n_chunks=iters, chunk_pad=(chunk + 127) // 128 * 128)
    if n_src_points < g.n_points:
        # the reference indexes points [0, 10000) unconditionally (csa_models.py:86-90)
        raise IndexError(f"index {g.n_points - 1} is out of bounds for dimension 2 with size {n_src_points}")
    return g


def _std_layout(t: torch.Tensor) -> torch.Tensor:
    """(.., 256, N, 1) fully contiguous (the kernels view it as a 2-D [channels of all shapes][N] matrix); copies
    only if needed."""
    return t if t.is_contiguous() else t.contiguous()


def _pack_sources(sources, n_slots, geom, dt, dev, want_f32: bool = True, want_sum: bool = False):
    """sources: list of (tensor (n0,[n1],256,N,1) fp32 cuda, slot0, d0, d1).  want_sum: also the per-chunk channel
    sums of every slot, [n_slots*n_chunks, 256] fp32 (the key means V is centred on)."""
    NP = geom.rows_pad
    Xh = torch.empty(n_slots * NP, 256, dtype=dt, device=dev)
    Xf = torch.empty(n_slots * NP, 256, dtype=torch.float32, device=dev) if want_f32 else None
    xsum = torch.zeros(n_slots * geom.n_chunks, 256, dtype=torch.float32, device=dev) if want_sum else None
    for (t, slot0, d0, d1) in sources:
        if t.dim() == 4:
            t = t.unsqueeze(1)
        assert t.dim() == 5 and t.shape[2] == 256 and t.shape[4] == 1, t.shape
        if t.stride(3) != 1 or t.stride(2) != t.shape[3]:
            t = t.contiguous()
        E._launch_pack(t, Xh, Xf, t.shape[0], t.stride(0), t.shape[1], t.stride(1), slot0, d0, d1, geom, t.shape[3], xsum)
    return (Xh, Xf, xsum) if want_sum else (Xh, Xf)


def _unpad_rows(Y: torch.Tensor, n_blocks: int, geom: E.Geometry) -> torch.Tensor:
    return Y.view(n_blocks, geom.n_chunks, geom.chunk_pad, 256)[:, :, :geom.chunk].reshape(n_blocks, geom.n_points, 256)


def _pad_rows(dY: torch.Tensor, geom: E.Geometry) -> torch.Tensor:
    nb = dY.shape[0]
    out = torch.zeros(nb, geom.n_chunks, geom.chunk_pad, 256, dtype=torch.float32, device=dY.device)
    out[:, :, :geom.chunk] = dY.reshape(nb, geom.n_chunks, geom.chunk, 256)
    return out.view(nb * geom.rows_pad, 256)


def _rows_to_channel_major(rows: torch.Tensor, n_b: int, n_src_points: int, geom: E.Geometry) -> torch.Tensor:
    """[n_b*NP,256] padded fp32 rows -> (n_b,256,n_src_points,1) (zeros beyond the used points)."""
    dev = rows.device
    alloc = torch.zeros if n_src_points > geom.n_points else torch.empty
    out = alloc(n_b, 256, n_src_points, 1, dtype=torch.float32, device=dev)
    blk = E.cached_table(("arange", n_b), dev, lambda: torch.arange(n_b, dtype=torch.int32))
    w = E.cached_table(("ones", n_b), dev, lambda: torch.ones(n_b, dtype=torch.float32))
    rc = L.lib().csn_combine_fwd(rows.data_ptr(), blk.data_ptr(), w.data_ptr(), out.data_ptr(), None, n_b, 1,
                                 256 * n_src_points, n_src_points, geom.n_points, geom.chunk, geom.chunk_pad,
                                 geom.rows_pad, L.CSN_F16, None, None, None, None, L.stream_ptr())
    L.check(rc, "csn_combine_fwd")
    return out


def _last_chunk_attn(ctx: E.AttnContext, blocks: int) -> torch.Tensor:
    """Attention matrix of the last chunk only (csa_models.py:125): (blocks, h, 500, 500)."""
    g = ctx.geom
    if ctx.P is None:
        # fused forward keeps no probabilities: rebuild the last chunk from Q, K (cheap: 1/20 of QK^T)
        return _attn_of_chunk(ctx, blocks, g.n_chunks - 1)
    P = ctx.P.view(ctx.n_blocks, g.n_chunks, ctx.n_head, g.chunk_pad, g.chunk_pad)
    return P[:blocks, -1, :, :g.chunk, :g.chunk].float()


def _attn_of_chunk(ctx: E.AttnContext, blocks: int, c: int) -> torch.Tensor:
    import math
    g, h, d = ctx.geom, ctx.n_head, ctx.d_head
    NP, CP = g.rows_pad, g.chunk_pad
    HD = h * d
    dev, dt = ctx.Xh.device, ctx.Xh.dtype
    table = {j: (qs, ks) for grp in ctx.groups for (j, qs, ks, _) in grp.blocks()}
    S = torch.empty(blocks * h * CP, CP, dtype=torch.float32, device=dev)
    Qv, Kv = ctx.QKV[:, :HD], ctx.QKV[:, HD:2 * HD]
    for j in range(blocks):
        qs, ks = table[j]
        A = L.mat(Qv[qs * NP + c * CP:], L.MAJOR_K, k_off=(d,))
        B = L.mat(Kv[ks * NP + c * CP:], L.MAJOR_K, k_off=(d,))
        L.gemm(A, B, L.out(S[j * h * CP:], CP, off=(CP * CP,)), CP, CP, d, nb=(h,), alpha=1.0 / math.sqrt(d))
    P = torch.empty(blocks * h * CP, CP, dtype=dt, device=dev)
    rc = L.lib().csn_softmax_fwd(S.data_ptr(), P.data_ptr(), S.shape[0], CP, g.chunk, CP, g.chunk, L.dtype_code(dt),
                                 L.stream_ptr())
    L.check(rc, "csn_softmax_fwd")
    return P.view(blocks, h, CP, CP)[:, :, :g.chunk, :g.chunk].float()


# ------------------------------------------------------------------------------------------- MHA op
class _MhaFn(torch.autograd.Function):
    """MultiHeadAttention.forward (csa_models.py:81-125) on channel-major inputs."""

    @staticmethod
    def forward(ctx, Q, K, V, wq, wk, wv, wo, gamma, beta, n_head, dt, iters, chunk, dropout_p=0.0, seed=0,
                same_qk=False, same_kv=False):
        for t, n in ((Q, "Q"), (K, "K"), (V, "V"), (wq, "w_qs.weight")):
            _require_cuda(t, n)
        B, n_src = Q.shape[0], Q.shape[2]
        geom = _geom_for(min(Q.shape[2], K.shape[2], V.shape[2]), iters, chunk)
        # same_qk / same_kv: the caller passed the SAME tensor object twice (decided by the module on object identity:
        # an alias such as k = q.detach() is a different autograd leaf and gets its own slot and its own gradient)
        sources = [(Q.float(), 0, 1, 0)]
        k0 = 0 if same_qk else B
        if not same_qk:
            sources.append((K.float(), B, 1, 0))
        n_slots = k0 + B
        v0 = k0
        if not same_kv:
            v0 = n_slots
            sources.append((V.float(), v0, 1, 0))
            n_slots += B
        Xh, Xf = _pack_sources(sources, n_slots, geom, dt, Q.device)
        group = E.Group(n_in=B, n_out=1, blk0=0, q0=0, q_si=1, q_so=0, k0=k0, k_si=1, k_so=0, v0=v0, v_si=1, v_so=0)
        a = E.attention_forward(Xh, Xf, [group], n_slots, B, wq, wk, wv, wo, gamma, beta, geom, n_head,
                                want_colsum=False, dropout_p=dropout_p, seed=seed)
        ctx.a = a
        ctx.meta = (B, k0, v0, n_slots, Q.shape[2], K.shape[2], V.shape[2])
        attn = _last_chunk_attn(a, B)
        ctx.mark_non_differentiable(attn)
        return _unpad_rows(a.Y, B, geom), attn

    @staticmethod
    def backward(ctx, dret, _dattn):
        a = ctx.a
        B, k0, v0, n_slots, nq, nk, nv = ctx.meta
        need_dx = any(ctx.needs_input_grad[:3])
        g = E.attention_backward(a, _pad_rows(dret.float().contiguous(), a.geom), need_dx)
        dQ = dK = dV = None
        if need_dx:
            NP = a.geom.rows_pad
            dX = g["dX"]
            dq_rows = dX[:B * NP]
            dQ = _rows_to_channel_major(dq_rows, B, nq, a.geom)
            if k0 == 0:      # Q is K (and possibly V): the gradient belongs to the one tensor
                dK = None
            else:
                dK = _rows_to_channel_major(dX[k0 * NP:(k0 + B) * NP], B, nk, a.geom)
            if v0 != k0:
                dV = _rows_to_channel_major(dX[v0 * NP:(v0 + B) * NP], B, nv, a.geom)
        return (dQ, dK, dV, g["dWq"], g["dWk"], g["dWv"], g["dWo"], g["dgamma"], g["dbeta"], None, None, None, None,
                None, None, None, None)


class ScaledDotProductAttention(nn.Module):
    """csa_models.py:128-144.  q,k,v: (B,h,L,d) fp32 CUDA -> (output (B,h,Lq,d), attn (B,h,Lq,Lk)).
    Standalone use only (the fused layer never calls it); forward pass, no autograd."""

    def __init__(self, temperature, attn_dropout=0.1, precision="fp16"):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)
        self.precision = precision

    def forward(self, q, k, v):
        _require_cuda(q, "q")
        dt = _PRECISIONS[self.precision]
        B, h, Lq, d = q.shape
        Lk = k.shape[2]
        Lkp = (Lk + 7) // 8 * 8
        q16 = q.reshape(B * h * Lq, d).to(dt)
        k16 = k.reshape(B * h * Lk, d).to(dt)
        v16 = v.reshape(B * h * Lk, d).to(dt)
        S = torch.empty(B * h * Lq, Lkp, dtype=torch.float32, device=q.device)
        L.gemm(L.mat(q16, L.MAJOR_K, mn_off=(Lq,)), L.mat(k16, L.MAJOR_K, mn_off=(Lk,)),
               L.out(S, Lkp, off=(Lq * Lkp,)), Lq, Lk, d, nb=(B * h,), alpha=1.0 / float(self.temperature))
        P = torch.empty(B * h * Lq, Lkp, dtype=dt, device=q.device)
        rc = L.lib().csn_softmax_fwd(S.data_ptr(), P.data_ptr(), S.shape[0], Lkp, Lk, 1, 1, L.dtype_code(dt),
                                     L.stream_ptr())
        L.check(rc, "csn_softmax_fwd")
        out = torch.empty(B * h * Lq, d, dtype=torch.float32, device=q.device)
        L.gemm(L.mat(P, L.MAJOR_K, mn_off=(Lq,)), L.mat(v16, L.MAJOR_MN, k_off=(Lk,)),
               L.out(out, d, off=(Lq * d,)), Lq, d, Lk, nb=(B * h,))
        return out.view(B, h, Lq, d), P[:, :Lk].float().view(B, h, Lq, Lk)


class MultiHeadAttention(nn.Module):
    """Multi-Head Attention module (csa_models.py:37-125): block-diagonal attention in 20 chunks of
    500 points + output projection + residual + LayerNorm(eps=1e-6)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, precision="fp16"):
        super().__init__()
        assert d_k == d_v, "the kernels assume d_k == d_v (true for every reference configuration)"
        self.n_head = n_head
        self.d_k = d_k
        self.d_v = d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.attention = ScaledDotProductAttention(temperature=d_k ** 0.5, precision=precision)
        self.dropout = nn.Dropout(dropout)
        self.norm = nn.LayerNorm(d_model, eps=1e-6)
        self.precision = precision
        self.iters = 20      # csa_models.py:83
        self.mini_bs = 500   # csa_models.py:84

    def _weights(self):
        return (self.w_qs.weight, self.w_ks.weight, self.w_vs.weight, self.fc.weight, self.norm.weight, self.norm.bias)

    def forward(self, Q, K, V, mode=None):
        """Q,K,V: (B,256,N,1) channel-major. Returns (ret (B,10000,256), attn of the last chunk
        (B,h,500,500)).  `mode` is ignored, as in the reference."""
        return _MhaFn.apply(Q, K, V, *self._weights(), self.n_head, _PRECISIONS[self.precision], self.iters,
                            self.mini_bs, *self._dropout_state(), Q is K, K is V)

    def _dropout_state(self):
        """(p, seed): dropout follows `module.training` (csa_models.py:56,136; the `mode` argument is ignored).
        One p serves both nn.Dropout modules of the reference (attention probabilities and fc output: both 0.1).  The
        seed comes from torch's CPU generator, so torch.manual_seed reproduces a run."""
        p = float(self.dropout.p) if self.training else 0.0
        if p <= 0.0:
            return 0.0, 0
        return p, int(torch.randint(0, 2 ** 31 - 1, (1,)).item())

    def self_attention(self, x):
        """csa_models.py:59-79 (unused by the reference): full, un-chunked self-attention."""
        n = x.shape[2]
        return _MhaFn.apply(x, x, x, *self._weights(), self.n_head, _PRECISIONS[self.precision], 1, n,
                            *self._dropout_state(), True, True)


# ------------------------------------------------------------------------------------------- fused segmentation loss
class _SegLossFn(torch.autograd.Function):
    """loss = masked cross-entropy(logit(feats), labels) (csa_models.py:201 + csa_training.py:94-108) with the
    logits never materialised on the forward path: one pass over the activation computes the loss, d loss / d feats
    and d loss / d logits (for the weight gradient)."""

    @staticmethod
    def _run(f, W, lab, ignore_index, want_grads, gout):
        B, D, N = f.shape
        Cn = W.shape[0]
        dev = f.device
        n_valid = torch.zeros(1, dtype=torch.int32, device=dev)
        parts = torch.empty(B * ((N + 127) // 128), dtype=torch.float32, device=dev)
        dlogits = torch.empty(B, Cn, N, dtype=torch.float32, device=dev) if want_grads else None
        dfeat = torch.empty_like(f) if want_grads else None
        rc = L.lib().csn_seg_loss(f.data_ptr(), D * N, N, B, N, W.data_ptr(), Cn, lab.data_ptr(), int(ignore_index),
                                  n_valid.data_ptr(), parts.data_ptr(), dlogits.data_ptr() if want_grads else None,
                                  dfeat.data_ptr() if want_grads else None, gout.data_ptr() if gout is not None else None,
                                  L.stream_ptr())
        L.check(rc, "csn_seg_loss")
        return parts, n_valid, dlogits, dfeat

    @staticmethod
    def forward(ctx, feats, weight, labels, ignore_index):
        _require_cuda(feats, "feats")
        B, D, N = feats.shape[0], feats.shape[1], feats.shape[2]
        assert D == 256
        f = feats.float().contiguous().view(B, D, N)
        Cn = weight.shape[0]
        W = weight.detach().float().reshape(Cn, D).contiguous()
        lab = labels.reshape(B, N).to(torch.int64).contiguous()
        parts, n_valid, _, _ = _SegLossFn._run(f, W, lab, ignore_index, False, None)
        ctx.save_for_backward(f, W, lab)
        ctx.meta = (feats.shape, weight.shape, ignore_index)
        return parts.sum() / n_valid.clamp_min(1).to(torch.float32)[0]

    @staticmethod
    def backward(ctx, gout):
        # the pass is repeated with the gradient outputs switched on and the upstream gradient as a device scalar:
        # recomputing 0.6 GFLOP of logits is cheaper than keeping (and rescaling) 82 MB of d feats from the forward
        f, W, lab = ctx.saved_tensors
        fshape, wshape, ignore_index = ctx.meta
        g = gout.detach().float().reshape(1).contiguous()
        _, _, dlogits, dfeat = _SegLossFn._run(f, W, lab, ignore_index, True, g)
        gf = dfeat.view(fshape) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:   # d logits f^T: B small fp32 library GEMMs on a transposed view (no copies)
            gw = torch.bmm(dlogits, f.transpose(1, 2)).sum(0).view(wshape)
        return gf, gw, None, None


def segmentation_loss(feats: torch.Tensor, logit_weight: torch.Tensor, labels: torch.Tensor, ignore_index: int = 0):
    """Masked cross-entropy of the MID-FC training scripts (csa_training.py:94-108) on top of the bias-free 1x1
    `logit` conv (csa_models.py:201), fused: feats (B,256,N,1), logit_weight (C,256,1,1), labels (B,N);
    points whose label == ignore_index (0 in the reference) are masked; mean over the others."""
    return _SegLossFn.apply(feats, logit_weight, labels, ignore_index)


# ------------------------------------------------------------------------------------------- CSA op
class _CsaState:
    """What the backward pass of the CSA layer needs from its forward pass (shared by _CsaFn and _CsaLossFn)."""
    __slots__ = ("a", "glue", "comp", "blk", "meta")


def _csa_forward_core(x, x_neighbors, wq, wk, wv, wo, gamma, beta, cq_w, cq_b, ck_w, ck_b, n_head, dt, iters, chunk,
                      dropout_p: float = 0.0, seed: int = 0) -> _CsaState:
    """Everything of get_csa_feats (csa_models.py:209-230) up to the compatibility weights: packing, projections,
    the (2K+1)B attention blocks, residual + LayerNorm statistics, pooled means, compatibility softmax."""
    _require_cuda(x, "x")
    _require_cuda(wq, "attention.w_qs.weight")
    dev = x.device
    B, n_src = x.shape[0], x.shape[2]
    ssa_only = x_neighbors is None
    K = 0 if ssa_only else x_neighbors.shape[1] - 1
    if not ssa_only and not x_neighbors.is_cuda:
        # the reference moves each neighbour inside the layer (csa_models.py:216,236)
        x_neighbors = neighbors_to_device(x_neighbors, dev)
    n_src_nb = n_src if ssa_only or K == 0 else x_neighbors.shape[3]
    geom = _geom_for(min(n_src, n_src_nb), iters, chunk)
    S = B * (K + 1)
    # 16-bit inputs (a caller-side 16-bit feature cache: half the host-to-device bytes) are packed as they are; the
    # residual then comes from the fp32 copy the pack kernel writes instead of the caller's channel-major tensors
    in16 = x.dtype in (torch.float16, torch.bfloat16) and (ssa_only or x_neighbors.dtype == x.dtype)
    xs = _std_layout(x if in16 else x.float())
    sources = [(xs, 0, K + 1, 0)]
    nbs = None
    if K > 0:
        nbs = _std_layout(x_neighbors if in16 else x_neighbors.float())
        sources.append((nbs[:, 1:], 1, K + 1, 1))
    # the residual of the output projection is read straight from these channel-major tensors by the fused
    # projection/LayerNorm kernel (slot s = b*(K+1)+j: j = 0 the query, j >= 1 neighbour j)
    res_cm = None
    # (TMA needs 16-byte aligned rows: point counts that are not multiples of 4 take the unfused path)
    tma_ok = (not in16) and xs.shape[2] % 4 == 0 and xs.data_ptr() % 16 == 0 and (nbs is None or nbs.data_ptr() % 16 == 0)
    if tma_ok and (K == 0 or nbs.shape[3] == xs.shape[2]):
        sel = tuple(0 if j == 0 else 1 for b in range(B) for j in range(K + 1))
        off = tuple(b * xs.stride(0) if j == 0 else b * nbs.stride(0) + j * nbs.stride(1)
                    for b in range(B) for j in range(K + 1))
        res_cm = E.ChannelMajorResidual(bases=(xs,) if K == 0 else (xs, nbs), sel=sel, off=off,
                                        ch_stride=xs.shape[2], n_points=xs.shape[2])
    Xh, Xf, xsum = _pack_sources(sources, S, geom, dt, dev, want_f32=res_cm is None or not E.use_fused_ln(),
                                 want_sum=True)
    groups = [E.Group(n_in=S, n_out=1, blk0=0, q0=0, q_si=1, q_so=0, k0=0, k_si=1, k_so=0, v0=0, v_si=1, v_so=0)]
    nblk = S
    if K > 0:
        groups.append(E.Group(n_in=K, n_out=B, blk0=S, q0=0, q_si=0, q_so=K + 1, k0=1, k_si=1, k_so=K + 1,
                              v0=1, v_si=1, v_so=K + 1))
        nblk += B * K
    a = E.attention_forward(Xh, Xf, groups, S, nblk, wq, wk, wv, wo, gamma, beta, geom, n_head,
                            want_colsum=not ssa_only, want_y=False, residual_cm=res_cm, colsum_blocks=S,
                            dropout_p=dropout_p, seed=seed, chunk_sum=xsum)
    # ---- compatibility (csa_models.py:211-230): csn_compat_fwd
    if ssa_only:
        comp = torch.ones(B, 1, dtype=torch.float32, device=dev)
        glue = None
    else:
        # csa_models.py:222-230 incl. the batch-interleaving view, fp64 inside (csrc/compat.cu): the backward
        # of this softmax subtracts d comp values that agree in their first 3-5 digits, and the tensors are tiny
        pooled = a.colsum[:S].contiguous()                              # slot order (b, k), fp32
        K1 = K + 1
        u_q = torch.empty(B, 256, dtype=torch.float64, device=dev)
        u_k = torch.empty(S, 256, dtype=torch.float64, device=dev)
        nrm = torch.empty(B + S, dtype=torch.float64, device=dev)
        comp64 = torch.empty(B, K1, dtype=torch.float64, device=dev)
        comp = torch.empty(B, K1, dtype=torch.float32, device=dev)
        wts = tuple(t.detach().float().contiguous() for t in (cq_w, cq_b, ck_w, ck_b))
        rc = L.lib().csn_compat_fwd(pooled.data_ptr(), wts[0].data_ptr(), wts[1].data_ptr(), wts[2].data_ptr(), wts[3].data_ptr(),
                                    B, K1, u_q.data_ptr(), u_k.data_ptr(), nrm[:B].data_ptr(), nrm[B:].data_ptr(),
                                    comp64.data_ptr(), comp.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_compat_fwd")
        glue = (pooled, wts, u_q, u_k, nrm, comp64)

    def _blk_table():
        t = torch.empty(B, K + 1, dtype=torch.int32)
        for b in range(B):
            t[b, 0] = b * (K + 1)
            for k in range(1, K + 1):
                t[b, k] = S + b * K + (k - 1)
        return t
    st = _CsaState()
    st.a, st.glue, st.comp = a, glue, comp.contiguous()
    st.blk = E.cached_table(("csa_blk", B, K), dev, _blk_table)
    st.meta = (B, K, S, nblk, n_src, n_src_nb)
    return st


def _csa_bwd_tables(st: _CsaState, dev):
    B, K, S, nblk, _, _ = st.meta
    has_glue = st.glue is not None

    def _bwd_tables():
        cb = torch.full((nblk,), -1, dtype=torch.int32)
        cwi = torch.full((nblk,), -1, dtype=torch.int32)
        pb = torch.full((nblk,), -1, dtype=torch.int32)
        for b in range(B):
            for k in range(K + 1):
                slot = b * (K + 1) + k
                if has_glue:
                    pb[slot] = slot
                if k == 0:
                    cb[slot], cwi[slot] = b, slot
                else:
                    j = S + b * K + (k - 1)
                    cb[j], cwi[j] = b, slot
        return torch.stack([cb, cwi, pb, torch.full((nblk,), -1, dtype=torch.int32)])
    tabs = E.cached_table(("csa_bwd", B, K, has_glue), dev, _bwd_tables)
    gidx = E.cached_table(("csa_bwd_gather", B, K, has_glue), dev, lambda: _bwd_tables()[1].clamp(min=0).long())
    return tabs[0], tabs[1], tabs[2], gidx


def _csa_backward_core(st: _CsaState, dOutT, amax, dcomp, need_dx: bool, out_scale=None):
    """Backward of _csa_forward_core + the weighted sum, given the output gradient as padded rows dOutT [B*NP, 256]
    (zero in pad rows), amax >= max|dOutT| (device scalar) and dcomp[b,k] = <dOut[b], MHA_k> (or None without glue).
    out_scale: optional device scalar multiplied into the whole upstream gradient (d loss of a fused head)."""
    a = st.a
    geom = a.geom
    B, K, S, nblk, n_src, n_src_nb = st.meta
    dev = dOutT.device
    cb, cwi, pb, gidx = _csa_bwd_tables(st, dev)
    grads_glue = [None] * 4
    dpool = None
    if st.glue is not None:
        pooled, wts, u_q, u_k, nrm, comp64 = st.glue
        K1 = K + 1
        dc = dcomp.reshape(-1)
        if dc.dtype != torch.float64:
            dc = dc.double()
        dlin = torch.empty(B + S, 256, dtype=torch.float64, device=dev)
        gW = torch.empty(2, 256, 256, dtype=torch.float32, device=dev)
        gb = torch.empty(2, 256, dtype=torch.float32, device=dev)
        dpool = torch.empty(S, 256, dtype=torch.float32, device=dev)
        dp_amax = torch.zeros(1, dtype=torch.float32, device=dev)
        rc = L.lib().csn_compat_bwd(pooled.data_ptr(), wts[0].data_ptr(), wts[2].data_ptr(), u_q.data_ptr(), u_k.data_ptr(),
                                    nrm[:B].data_ptr(), nrm[B:].data_ptr(), comp64.data_ptr(), dc.data_ptr(),
                                    out_scale.data_ptr() if out_scale is not None else None, B, K1,
                                    dlin[:B].data_ptr(), dlin[B:].data_ptr(), gW[0].data_ptr(), gb[0].data_ptr(),
                                    gW[1].data_ptr(), gb[1].data_ptr(), dpool.data_ptr(), dp_amax.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_compat_bwd")
        grads_glue = [gW[0], gb[0], gW[1], gb[1]]
    # per-block fan-out weights cw[j] = comp[b,k] * d loss, and the bound on |dY + dpool/N| for the gradient scaling
    cw = torch.empty(nblk, dtype=torch.float32, device=dev)
    amax_b = torch.empty(1, dtype=torch.float32, device=dev)
    rc = L.lib().csn_compat_fanout(st.comp.data_ptr(), cwi.data_ptr(), nblk, out_scale.data_ptr() if out_scale is not None else None,
                                   amax.data_ptr(), dp_amax.data_ptr() if st.glue is not None else None, 1.0 / geom.n_points,
                                   cw.data_ptr(), amax_b.data_ptr(), L.stream_ptr())
    L.check(rc, "csn_compat_fanout")
    amax = amax_b
    # The upstream gradient of block j is comp[b,k] * dOut[b]^T (+ the pooled-mean row vector): csn_ln_bwd forms
    # cw[j] * dOutT[cb[j]] + dpool[pb[j]]/N on the fly; the (2K+1)x larger dY is never materialised.
    g = E.attention_backward(a, dOutT, need_dx, amax, bcast=dpool, bcast_idx=pb if dpool is not None else None,
                             bcast_scale=1.0 / geom.n_points, src_idx=cb, src_w=cw)
    return g, grads_glue


class _CsaFn(torch.autograd.Function):
    """CrossShapeAt.get_csa_feats (csa_models.py:209-242) / get_ssa_feats (:204-207, x_neighbors=None)."""

    @staticmethod
    def forward(ctx, x, x_neighbors, wq, wk, wv, wo, gamma, beta, cq_w, cq_b, ck_w, ck_b, n_head, dt, iters, chunk,
                want_attn=True, dropout_p=0.0, seed=0):
        st = _csa_forward_core(x, x_neighbors, wq, wk, wv, wo, gamma, beta, cq_w, cq_b, ck_w, ck_b, n_head, dt, iters,
                               chunk, dropout_p, seed)
        a, geom = st.a, st.a.geom
        B, K, S, nblk, n_src, n_src_nb = st.meta
        dev = x.device
        # ---- out = sum_k comp[b,k] * MHA(x, x_k)   (:232-240), written channel-major
        alloc = torch.zeros if n_src > geom.n_points else torch.empty
        out = alloc(B, 256, geom.n_points, 1, dtype=torch.float32, device=dev)
        # the LayerNorm output is formed on the fly from z, mean, rstd (it is never written to HBM)
        rc = L.lib().csn_combine_fwd(a.Z.data_ptr(), st.blk.data_ptr(), st.comp.data_ptr(), out.data_ptr(), None, B, K + 1,
                                     256 * geom.n_points, geom.n_points, geom.n_points, geom.chunk, geom.chunk_pad,
                                     geom.rows_pad, L.dtype_code(dt), a.mean.data_ptr(), a.rstd.data_ptr(),
                                     gamma.data_ptr(), beta.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_combine_fwd")
        ctx.st = st
        if want_attn:
            attn = _last_chunk_attn(a, S).view(B, K + 1, n_head, geom.chunk, geom.chunk)[:, 0]
        else:   # every reference caller discards it
            attn = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, dout, _dattn):
        st = ctx.st
        a, geom = st.a, st.a.geom
        B, K, S, nblk, n_src, n_src_nb = st.meta
        dev = dout.device
        dout = dout.float().contiguous()
        has_glue = st.glue is not None
        cb, cwi, pb, gidx = _csa_bwd_tables(st, dev)
        lib = L.lib()
        # dOut is transposed ONCE per batch item into padded rows (it stays L2-resident)
        dOutT = torch.empty(B * geom.rows_pad, 256, dtype=torch.float32, device=dev)
        amax = torch.zeros(1, dtype=torch.float32, device=dev)
        rc = lib.csn_pack_rows(dout.data_ptr(), None, dOutT.data_ptr(), geom.n_points, B, 256 * geom.n_points, 1, 0, 0, 1, 0,
                               geom.n_points, geom.chunk, geom.chunk_pad, geom.rows_pad, L.CSN_F16, amax.data_ptr(),
                               None, L.stream_ptr())
        L.check(rc, "csn_pack_rows(dOut)")
        # d comp[b,k] = <dOut[b]^T, MHA_k> with the LayerNorm output re-formed from z on the fly
        dcomp = None
        if has_glue:
            dcomp = torch.zeros(B * (K + 1), dtype=torch.float32, device=dev)
            rc = lib.csn_block_dot(dOutT.data_ptr(), a.Z.data_ptr(), cb.data_ptr(), cwi.data_ptr(), dcomp.data_ptr(),
                                   nblk * geom.rows_pad, geom.rows_pad, geom.chunk_pad, geom.chunk, a.mean.data_ptr(),
                                   a.rstd.data_ptr(), a.gamma.data_ptr(), a.beta.data_ptr(), L.stream_ptr())
            L.check(rc, "csn_block_dot")
        need_dx = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g, grads_glue = _csa_backward_core(st, dOutT, amax, dcomp, need_dx)
        dx, dnb = _csa_input_grads(g, st, ctx.needs_input_grad[0], ctx.needs_input_grad[1]) if need_dx else (None, None)
        return (dx, dnb, g["dWq"], g["dWk"], g["dWv"], g["dWo"], g["dgamma"], g["dbeta"], *grads_glue,
                None, None, None, None, None, None, None)


def _csa_input_grads(g, st: _CsaState, want_x: bool, want_nb: bool):
    B, K, S, nblk, n_src, n_src_nb = st.meta
    geom = st.a.geom
    G = _rows_to_channel_major(g["dX"], S, max(n_src, n_src_nb), geom).view(B, K + 1, 256, -1, 1)
    dx = G[:, 0, :, :n_src].contiguous() if want_x else None
    dnb = None
    if want_nb:
        dnb = G[:, :, :, :n_src_nb].clone()
        dnb[:, 0] = 0   # slot 0 of x_neighbors is never read (csa_models.py:214,234)
    return dx, dnb


class _CsaLossFn(torch.autograd.Function):
    """forward_csa / forward_ssa + the training scripts' masked cross-entropy (csa_models.py:182-242 +
    csa_training.py:94-134) with the fused head kernel csn_csa_head: the weighted sum of the K+1 attention outputs, the
    logit conv, the loss, the IoU counters AND their backward (d out as padded rows, d comp, d logit.weight) are one
    pass over the pre-LayerNorm rows.  Returns (loss, stats) with stats = int32 [3*C + 2]:
    #pred==c, #label==c, #(pred==c & label==c) over the unmasked points, then #correct and #labels out of range."""

    @staticmethod
    def forward(ctx, x, x_neighbors, labels, logit_w, ignore_index, wq, wk, wv, wo, gamma, beta, cq_w, cq_b, ck_w, ck_b,
                n_head, dt, iters, chunk, dropout_p=0.0, seed=0):
        st = _csa_forward_core(x, x_neighbors, wq, wk, wv, wo, gamma, beta, cq_w, cq_b, ck_w, ck_b, n_head, dt, iters,
                               chunk, dropout_p, seed)
        a, geom = st.a, st.a.geom
        B, K, S, nblk, n_src, n_src_nb = st.meta
        dev = x.device
        Cn = logit_w.shape[0]
        W = logit_w.detach().float().reshape(Cn, 256).contiguous()
        lab = labels.reshape(B, -1)
        if lab.dtype != torch.int64 or lab.stride(1) != 1:
            lab = lab.to(torch.int64).contiguous()
        if lab.shape[1] < geom.n_points:
            raise IndexError(f"labels hold {lab.shape[1]} points per shape, the layer uses {geom.n_points}")
        want_grad = any(ctx.needs_input_grad)
        lib = L.lib()
        NP = geom.rows_pad
        grid = lib.csn_csa_head_grid(B, NP)   # one persistent CTA per SM: per-CTA partial sums
        # zero-initialised scratch: [n_valid | stats (3C+2)], [amax | loss partials], d comp (fp64)
        ints = torch.zeros(1 + 3 * Cn + 2, dtype=torch.int32, device=dev)
        flts = torch.zeros(2 + grid, dtype=torch.float32, device=dev)
        n_valid, stats = ints[:1], ints[1:]
        amax, loss_buf, loss_part = flts[:1], flts[1:2], flts[2:]
        dcomp = torch.zeros(B * (K + 1), dtype=torch.float64, device=dev)
        dOutT = torch.empty(B * NP, 256, dtype=torch.float32, device=dev) if want_grad else None
        dW_part = torch.empty(grid, Cn, 256, dtype=torch.float32, device=dev) if want_grad else None
        dW = torch.empty(Cn, 256, dtype=torch.float32, device=dev) if want_grad else None
        has_glue = st.glue is not None
        rc = lib.csn_csa_head(a.Z.data_ptr(), a.mean.data_ptr(), a.rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                              st.blk.data_ptr(), st.comp.data_ptr(), B, K + 1, W.data_ptr(), Cn, lab.data_ptr(), lab.stride(0),
                              int(ignore_index), n_valid.data_ptr(), geom.n_points, geom.chunk, geom.chunk_pad, NP,
                              loss_part.data_ptr(), dOutT.data_ptr() if want_grad else None,
                              amax.data_ptr() if want_grad else None,
                              dcomp.data_ptr() if (want_grad and has_glue) else None,
                              dW_part.data_ptr() if want_grad else None, dW.data_ptr() if want_grad else None,
                              stats.data_ptr(), None, loss_buf.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_csa_head")
        loss = loss_buf[0]
        ctx.st = st
        ctx.saved = (dOutT, amax, dcomp if has_glue else None, dW, logit_w.shape)
        ctx.mark_non_differentiable(stats)
        return loss, stats

    @staticmethod
    def backward(ctx, gout, _gstats):
        st = ctx.st
        dOutT, amax, dcomp, dW, wshape = ctx.saved
        if dOutT is None:
            raise L.CsnError("forward_loss ran without gradients enabled: nothing to backpropagate")
        gs = gout.detach().float().reshape(1)
        need_dx = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g, grads_glue = _csa_backward_core(st, dOutT, amax, dcomp, need_dx, out_scale=gs)
        dx, dnb = _csa_input_grads(g, st, ctx.needs_input_grad[0], ctx.needs_input_grad[1]) if need_dx else (None, None)
        gw = (dW * gs).view(wshape) if ctx.needs_input_grad[3] else None
        return (dx, dnb, None, gw, None, g["dWq"], g["dWk"], g["dWv"], g["dWo"], g["dgamma"], g["dbeta"], *grads_glue,
                None, None, None, None, None, None)


class CrossShapeAt(nn.Module):
    """csa_models.py:146-404 with the same parameters / buffers / state_dict keys."""

    def __init__(self, num_classes, d_model, n_heads, K=None, d_k=256, d_v=256, attention_type='ssa', after_fc=False,
                 device=None, precision="fp16"):
        super().__init__()
        self.fc_1 = self.octree_conv1x1_bn_relu(928, 256)   # constructed but unused by forward (:150,:191-202)
        self.logit = self.octree_conv1x1(256, num_classes)
        self.attention = MultiHeadAttention(n_heads, d_model, d_k, d_v, precision=precision)
        self.attention_type = attention_type
        self.after_fc = after_fc
        self.device = device
        self.precision = precision
        if 'csa' in self.attention_type:
            self.K = K
            self.compatibility_q = nn.Linear(256, 256)
            self.compatibility_k = nn.Linear(256, 256)

    def octree_conv1x1_bn_relu(self, nin, nout):
        return nn.Sequential(self.octree_conv1x1_bn(nin, nout), nn.ReLU())

    def octree_conv1x1_bn(self, nin, nout):
        return nn.Sequential(self.octree_conv1x1(nin, nout, use_bias=False), nn.BatchNorm2d(nout))

    def octree_conv1x1(self, nin, nout, use_bias=False):
        layer = nn.Conv2d(nin, nout, kernel_size=1, stride=1, padding='same', bias=use_bias)
        nn.init.xavier_uniform_(layer.weight)
        return layer

    # -- forward paths (csa_models.py:182-202)
    def forward(self, x, mode, neighbor_feats=None):
        if self.attention_type == 'ssa':
            x = self.forward_ssa(x, mode)
        if self.attention_type == 'csa':
            x = self.forward_csa(x, neighbor_feats, mode)
        if self.attention_type == 'fcf_csaf_logitf':
            x = self.forward_fcf_csaf_logitf(x, neighbor_feats, mode)  # undefined in the reference too
        return x

    def forward_loss(self, x, mode, neighbor_feats, labels, ignore_index: int = 0, return_stats: bool = False):
        """forward(...) followed by the training scripts' masked cross-entropy (csa_training.py:94-108), with the weighted
        sum of the attention outputs, the logit conv, the loss, the accuracy / IoU counters (csa_training.py:110-134) and
        their backward fused into one pass over the attention blocks' pre-LayerNorm rows (csn_csa_head).
        return_stats=True also returns an int32 vector [3*C + 2]: per class #pred, #label, #(pred & label) over the
        unmasked points (intsc = the third, union = first + second - third), then #correct and #labels out of range."""
        if not self.after_fc:
            raise L.CsnError("forward_loss: the fused head follows the attention layer (after_fc=True models)")
        nb = neighbor_feats if self.attention_type == 'csa' else None
        p, seed = self._dropout_state()
        loss, stats = _CsaLossFn.apply(x, nb, labels, self.logit.weight, ignore_index, *self._csa_args(), p, seed)
        return (loss, stats) if return_stats else loss

    def _dropout_state(self):
        """(p, seed) of this call (see MultiHeadAttention._dropout_state).  Deviation from the reference, documented in
        DESIGN.md: SSA(x) is evaluated once per step and its single dropout mask serves both the pooled descriptor and
        the weighted sum (the reference calls it twice with independent masks, csa_models.py:210 vs :232)."""
        return self.attention._dropout_state() if self.training else (0.0, 0)

    def forward_ssa(self, x, mode):
        if self.after_fc:
            x, self_att = self.get_ssa_feats(x, mode)
        return self.logit(x)

    def forward_csa(self, x, x_neighbors, mode):
        if self.after_fc:
            x = self.get_csa_feats(x, x_neighbors, mode)
        return self.logit(x)

    def _csa_args(self):
        a = self.attention
        extra = (None, None, None, None)
        if 'csa' in self.attention_type:
            extra = (self.compatibility_q.weight, self.compatibility_q.bias, self.compatibility_k.weight,
                     self.compatibility_k.bias)
        return (*a._weights(), *extra, a.n_head, _PRECISIONS[self.precision], a.iters, a.mini_bs)

    def get_ssa_feats(self, x, mode):
        """(B,256,N,1) -> (SSA features (B,256,10000,1), attention of the last chunk)."""
        return _CsaFn.apply(x, None, *self._csa_args(), True, *self._dropout_state())

    def get_csa_feats(self, x, x_neighbors, mode):
        """x (B,256,N,1); x_neighbors (B,K+1,256,N,1) (slot 0 = the query, skipped; CPU tensors are
        accepted and moved, like csa_models.py:216,236) -> (B,256,10000,1)."""
        return _CsaFn.apply(x, x_neighbors, *self._csa_args(), False, *self._dropout_state())[0]

    # -- retrieval (csa_models.py:244-280)
    def get_retrieval_measure(self, ssa_feats_1, ssa_feats_2):
        return _knn.retrieval_measure(ssa_feats_1, ssa_feats_2, _PRECISIONS[self.precision])

    def get_knn_graph(self, ssa_feats_1, ssa_feats_2, K):
        return _knn.knn_graph(ssa_feats_1, ssa_feats_2, K, _PRECISIONS[self.precision])

    def get_all_feats(self, logs_dir, train_dataloader, K, mode):
        """csa_models.py:282-300: SSA features of every shape of a loader, (S, N, 256) on the CPU."""
        chunks = []
        dev = next(self.parameters()).device
        for feats, _label in train_dataloader:
            feats = torch.squeeze(feats.to(dev), dim=1)
            with torch.no_grad():
                batch, _ = self.get_ssa_feats(feats, mode)
            chunks.append(batch.detach().cpu())
        ssa = torch.cat(chunks, dim=0)
        return torch.permute(torch.squeeze(ssa, dim=-1), (0, 2, 1))

    def get_center_shape_indices(self, train_loader):
        """csa_models.py:302-332: k-means (S//10 centres) over amax-pooled SSA features; returns the
        shape nearest to each centre. KMeans itself is scikit-learn's, as in the reference."""
        from sklearn.cluster import KMeans
        dev = next(self.parameters()).device
        glob = []
        for feats, _label in train_loader:
            feats = torch.squeeze(feats.to(dev), dim=1)
            with torch.no_grad():
                batch, _ = self.get_ssa_feats(feats, 'test')
            glob.append(torch.amax(batch.squeeze(-1), dim=2).detach())
        glob = torch.cat(glob, dim=0).cpu().numpy()
        n_centers = len(glob) // 10
        kmeans = KMeans(n_clusters=n_centers, random_state=0, n_init=10).fit(glob)
        d = ((np.expand_dims(kmeans.cluster_centers_, 1) - glob) ** 2).sum(-1)
        return np.argmin(d, axis=-1)

    def get_candidate_ssa_feats(self, data_loader, candidate_shape_indices):
        """csa_models.py:334-358 (loader with batch size 1; indices sorted ascending)."""
        dev = next(self.parameters()).device
        out, counter = [], 0
        for i, (feats, _label) in enumerate(data_loader):
            if i != candidate_shape_indices[counter]:
                continue
            feats = torch.squeeze(feats.to(dev), dim=1)
            with torch.no_grad():
                batch, _ = self.get_ssa_feats(feats, 'test')
            out.append(batch.detach())
            counter += 1
            if counter == len(candidate_shape_indices):
                break
        ssa = torch.cat(out, dim=0)
        return torch.permute(torch.squeeze(ssa, dim=-1), (0, 2, 1)).contiguous()

    def get_retrieval_measure_big(self, query_loader, candidate_loader, candidate_shape_indices):
        """csa_models.py:360-392: every query of a loader against the candidate subset."""
        candidate_shape_indices.sort()
        cand = self.get_candidate_ssa_feats(candidate_loader, candidate_shape_indices)
        dt = _PRECISIONS[self.precision]
        cstore = _knn.build_store(cand, dt)
        dev = cand.device
        rows = []
        for feats, _label in query_loader:
            feats = torch.squeeze(feats.to(dev), dim=1)
            with torch.no_grad():
                f1, _ = self.get_ssa_feats(feats, 'test')
            f1 = f1.squeeze(-1).permute(0, 2, 1).contiguous()
            rows.append(_knn.scores_from_stores(_knn.build_store(f1, dt), cstore))
        return torch.cat(rows, dim=0)

    def get_knn_graph_big(self, query_loader, candidate_loader, candidate_shape_indices, K):
        s = self.get_retrieval_measure_big(query_loader, candidate_loader, candidate_shape_indices)
        return _knn.topk_rows(s.contiguous(), K + 1)[1]


def backbone_ssa_fc_logit(num_classes, n_heads, **kw):
    return CrossShapeAt(num_classes, 928, n_heads, attention_type='ssa', after_fc=False, **kw)


def backbone_fc_ssa_logit(num_classes, n_heads, **kw):
    return CrossShapeAt(num_classes, 256, n_heads, attention_type='ssa', after_fc=True, **kw)


def backbone_csa_fc_logit(num_classes, n_heads, K, **kw):
    return CrossShapeAt(num_classes, 928, n_heads, K, attention_type='csa', after_fc=False, **kw)


def backbone_fc_csa_logit(num_classes, n_heads, K, **kw):
    return CrossShapeAt(num_classes, 256, n_heads, K, attention_type='csa', after_fc=True, **kw)


def get_model(attention_type, num_classes, n_heads, K=None, **kw):
    """csa_models.py:426-432."""
    if attention_type == 'ssa':
        return backbone_fc_ssa_logit(num_classes, n_heads, **kw)
    elif attention_type == 'csa':
        return backbone_fc_csa_logit(num_classes, n_heads, K, **kw)
    raise AttributeError(f'{attention_type} not supported')
