"""Kernel orchestration for block attention (host side, no arithmetic).

An *attention block* is one independent MHA problem: the rows of a query shape attending to the rows
of a key/value shape, restricted chunk-by-chunk (MID-FC: 20 chunks of 500 points,
MID-FC/csa_models.py:83-90).  Shapes live in *slots* of padded row-major buffers
(slot s = rows [s*NP, (s+1)*NP) ); blocks are organised in regular *groups* so that one strided
batched launch covers a whole group.  Every contraction goes through csn_gemm (tcgen05), everything
else through the kernels of csrc/elementwise.cu; this file only computes strides.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import ctypes as C
import torch

from . import _lib as L


@dataclass(frozen=True)
class Geometry:
    """Chunked, padded row layout of one shape."""
    chunk: int = 500       # query points per attention chunk (mini_bs, csa_models.py:84)
    n_chunks: int = 20     # iters (csa_models.py:83)
    chunk_pad: int = 512   # rows reserved per chunk (multiple of 128)
    kv_chunk: int = -1     # key points per chunk when it differs from `chunk` (MinkowskiNet: Lq != Lk)

    @property
    def kv_len(self) -> int:
        return self.chunk if self.kv_chunk < 0 else self.kv_chunk

    @property
    def n_points(self) -> int:
        return self.chunk * self.n_chunks

    @property
    def rows_pad(self) -> int:
        return self.chunk_pad * self.n_chunks


@dataclass(frozen=True)
class Group:
    """n_out x n_in blocks; block (o, i) has index blk0 + o*n_in + i, query slot q0 + o*q_so + i*q_si,
    key slot k0 + o*k_so + i*k_si, value slot v0 + o*v_so + i*v_si."""
    n_in: int
    n_out: int
    blk0: int
    q0: int
    q_si: int
    q_so: int
    k0: int
    k_si: int
    k_so: int
    v0: int
    v_si: int
    v_so: int
    # ragged batches (MinkowskiNet: every shape has its own number of points): per block, in blocks() order, the
    # number of real query rows / key rows; empty = the geometry's chunk / kv_len for every block
    q_lens: tuple = ()
    kv_lens: tuple = ()

    def blocks(self):
        for o in range(self.n_out):
            for i in range(self.n_in):
                yield (self.blk0 + o * self.n_in + i, self.q0 + o * self.q_so + i * self.q_si,
                       self.k0 + o * self.k_so + i * self.k_si, self.v0 + o * self.v_so + i * self.v_si)


def _launch_pack(src: torch.Tensor, dst16, dst32, n0, s0, n1, s1, slot0, d0, d1, geom: Geometry, n_src_points,
                 chunk_sum=None):
    """src: fp32 (or fp16 / bf16) view whose element (i0,i1,c,n) sits at i0*s0 + i1*s1 + c*n_src_points + n.
    chunk_sum (optional, zero-initialised [n_slots*n_chunks, 256]): per-chunk channel sums of every packed shape."""
    tail = (dst16.data_ptr(), dst32.data_ptr() if dst32 is not None else None,
            n_src_points, n0, s0, n1, s1, slot0, d0, d1, geom.n_points, geom.chunk,
            geom.chunk_pad, geom.rows_pad, L.dtype_code(dst16.dtype), None,
            chunk_sum.data_ptr() if chunk_sum is not None else None, L.stream_ptr())
    if src.dtype in (torch.float16, torch.bfloat16):   # 16-bit feature cache (strides are in elements either way)
        rc = L.lib().csn_pack_rows_src16(src.data_ptr(), L.dtype_code(src.dtype), *tail)
    else:
        rc = L.lib().csn_pack_rows(src.data_ptr(), *tail)
    L.check(rc, "csn_pack_rows")


def sgemm_small(A, B, D, M, N, K, alpha=1.0, a_rows=None, b_rows=None, trans_a=False, trans_b=False, accumulate=False):
    """D[m][n] (+)= alpha * sum_k opA(m,k) opB(n,k) in fp32 on CUDA cores (csn_sgemm_small); 2-D row-major tensors."""
    rc = L.lib().csn_sgemm_small(A.data_ptr(), A.stride(0), a_rows.data_ptr() if a_rows is not None else None, int(trans_a),
                                 B.data_ptr(), B.stride(0), b_rows.data_ptr() if b_rows is not None else None, int(trans_b),
                                 D.data_ptr(), D.stride(0), M, N, K, float(alpha), int(accumulate), L.stream_ptr())
    L.check(rc, "csn_sgemm_small")


def use_centered_v() -> bool:
    """CSN_CENTER_V=1 centres V on its per-chunk key mean on the fused CSA path (no rounding residual of O, 16-bit
    O / P / V errors ~10x smaller: compatibility gradients within 1e-3 of fp64).  Off by default: measured on the
    config-2 step it saves 60 us (attention forward 403 -> 358 us, dO/delta GEMM 208 -> 193 us) and costs ~250 us
    (bias epilogues of the two projections, per-chunk sums in csn_ln_bwd, three small fp32 GEMMs); the default path
    already meets the fp64-calibrated gate (tests/test_midfc_gpu.py::test_compatibility_gradients_against_fp64)."""
    return os.environ.get("CSN_CENTER_V", "0") == "1"


def _kv_chunk_table(groups, n_blocks, n_chunks):
    """Row of the per-(slot, chunk) table that belongs to (block j, chunk c): key/value slot of j times n_chunks + c."""
    t = torch.empty(n_blocks, n_chunks, dtype=torch.int32)
    ar = torch.arange(n_chunks, dtype=torch.int32)
    for g in groups:
        for (j, _, ks, _) in g.blocks():
            t[j] = ks * n_chunks + ar
    return t.reshape(-1)


@dataclass
class AttnContext:
    """Everything the backward pass needs from the forward pass of a set of blocks."""
    geom: Geometry
    n_head: int
    d_head: int
    groups: list
    n_slots: int
    n_blocks: int
    Xh: torch.Tensor = None      # [S*NP, 256] 16-bit
    Xf: torch.Tensor = None      # [S*NP, 256] fp32
    QKV: torch.Tensor = None     # [S*NP, 3*h*d] 16-bit
    P: torch.Tensor = None       # [blocks*n_chunks*h*CP, CP] 16-bit probabilities
    O: torch.Tensor = None       # [blocks*NP, h*d] 16-bit
    Z: torch.Tensor = None       # [blocks*NP, 256] fp32 pre-LayerNorm
    Y: torch.Tensor = None       # [blocks*NP, 256] fp32 LayerNorm output
    mean: torch.Tensor = None
    rstd: torch.Tensor = None
    colsum: torch.Tensor = None  # [blocks, 256] mean of Y over the valid rows of each block
    Wqkv16: torch.Tensor = None
    Wo16: torch.Tensor = None
    gamma: torch.Tensor = None
    beta: torch.Tensor = None
    res_block: torch.Tensor = None
    extra: dict = field(default_factory=dict)


# Work tables live on the device and are built once per configuration.  Eviction is least-recently-used, one entry at
# a time; a captured CUDA graph holds raw pointers to the tables of its step, so csn_b200.graphs.GraphedStep pins the
# entries it saw (pin_tables) for as long as the graph lives.
from collections import OrderedDict

_ITEM_CACHE: "OrderedDict" = OrderedDict()
_TABLE_CACHE: "OrderedDict" = OrderedDict()
_ITEM_CACHE_MAX, _TABLE_CACHE_MAX = 64, 512


def pin_tables() -> list:
    """References to every cached work table (kept by a CUDA graph: evicted entries then stay allocated)."""
    return list(_ITEM_CACHE.values()) + list(_TABLE_CACHE.values())


def cached_table(key, device, build):
    """Small int/float index tables live on the device and are built once per configuration: a
    host->device copy from pageable memory inside a step would serialise the CPU with the stream."""
    k = (key, str(device))
    t = _TABLE_CACHE.get(k)
    if t is None:
        while len(_TABLE_CACHE) >= _TABLE_CACHE_MAX:
            _TABLE_CACHE.popitem(last=False)
        t = build().to(device)
        _TABLE_CACHE[k] = t
    else:
        _TABLE_CACHE.move_to_end(k)
    return t


def _res_block_table(groups, n_blocks):
    res = torch.empty(n_blocks, dtype=torch.int32)
    for g in groups:
        for (j, qs, _, _) in g.blocks():
            res[j] = qs
    return res


def attn_items(groups, geom: Geometry, n_head: int, d: int, device, kind: str = "fwd") -> torch.Tensor:
    """Work tables of the fused attention kernels for chunked block attention: one item per
    (block, chunk, head, 128-row tile). kind: 'fwd' (csn_attn_fwd), 'dv' (csn_attn_bwd_dv: tile of keys),
    'dq' (csn_attn_bwd_dq: tile of queries)."""
    key = (tuple(groups), geom, n_head, d, str(device), kind)
    t = _ITEM_CACHE.get(key)
    if t is not None:
        _ITEM_CACHE.move_to_end(key)
        return t
    NP, CP, NC = geom.rows_pad, geom.chunk_pad, geom.n_chunks
    tiles = CP // 128
    rows = []
    for g in groups:
        for bi, (j, qs, ks, vs) in enumerate(g.blocks()):
            assert ks == vs, "fused attention reads K and V rows from the same slot"
            q_len = g.q_lens[bi] if g.q_lens else geom.chunk       # real query rows per chunk of this block
            kv_len = g.kv_lens[bi] if g.kv_lens else geom.kv_len   # real key rows
            ragged = bool(g.q_lens or g.kv_lens)   # tiles that lie entirely in the padding are skipped: the caller
                                                    # zero-initialises O / dQ|dK|dV / dS for ragged batches
            for c in range(NC):
                for h in range(n_head):
                    for t_ in range(tiles):
                        valid = max(0, min(128, q_len - t_ * 128))      # query rows of tile t_
                        kvalid = max(0, min(128, kv_len - t_ * 128))    # key rows of tile t_
                        stat = (j * n_head + h) * NP + c * CP
                        if ragged and (kvalid if kind in ("dv", "dkv") else valid) == 0:
                            continue
                        if kind == "fwd":
                            rows.append((qs * NP + c * CP + t_ * 128, valid, ks * NP + c * CP, kv_len,
                                         j * NP + c * CP + t_ * 128, h * d, stat + t_ * 128, 1, ks * NP + c * CP, 0))
                        elif kind == "dv":   # resident tile = keys; streamed = queries (Q view) and dO (block rows)
                            rows.append((ks * NP + c * CP + t_ * 128, kvalid, qs * NP + c * CP, q_len,
                                         j * NP + c * CP + t_ * 128, h * d, stat, 1, j * NP + c * CP, t_ * 128))
                        elif kind == "dkv":  # csn_attn_bwd_dkv: resident keys, streamed queries + dO, per-query statistics
                            rows.append((ks * NP + c * CP + t_ * 128, kvalid, qs * NP + c * CP, q_len,
                                         j * NP + c * CP + t_ * 128, h * d, stat, j * NP + c * CP, t_ * 128, 0, 0, 0))
                        else:                # dq
                            rows.append((qs * NP + c * CP + t_ * 128, valid, ks * NP + c * CP, kv_len,
                                         j * NP + c * CP + t_ * 128, h * d, stat + t_ * 128,
                                         ((j * NC + c) * n_head + h) * CP + t_ * 128, 0, 1, 0, 0))
    if d != 256 and any(g.q_lens or g.kv_lens for g in groups) and os.environ.get("CSN_ITEM_SORT", "1") == "1":
        # ragged batch: items cost ~ their streamed length (column 3 of every item layout).  The persistent CTAs take
        # items round-robin, so handing them out longest-first keeps the per-CTA sums within ~1.5 % of each other
        # (6 % in table order on the config-4 batch).  (d_head 256 pairs items 2m / 2m+1 into clusters: order kept.)
        rows.sort(key=lambda r: -r[3])
    t = torch.tensor(rows, dtype=torch.int32).to(device)
    while len(_ITEM_CACHE) >= _ITEM_CACHE_MAX:
        _ITEM_CACHE.popitem(last=False)
    _ITEM_CACHE[key] = t
    return t


def _paired(geom: Geometry) -> int:
    """Items are emitted tile-fastest, so items 2m / 2m+1 are two 128-row tiles of the same (block,
    chunk, head) — they stream identical tiles — whenever a chunk has an even number of tiles."""
    return int((geom.chunk_pad // 128) % 2 == 0 and os.environ.get("CSN_CLUSTER", "1") != "0")


def use_fused_attention(d: int) -> bool:
    import os
    return d in (64, 256) and os.environ.get("CSN_FUSED_ATTN", "1") != "0"


def ctx_with(ctx, QKV):
    ctx.QKV = QKV
    return ctx


def _scores_and_probs(ctx: "AttnContext"):
    """Materialised path (also used by backward to rebuild P): S = QK^T/sqrt(d) -> softmax -> 16-bit P."""
    geom, n_head, d = ctx.geom, ctx.n_head, ctx.d_head
    NP, CP, NC = geom.rows_pad, geom.chunk_pad, geom.n_chunks
    HD = n_head * d
    dev, dt = ctx.Xh.device, ctx.Xh.dtype
    Qv, Kv = ctx.QKV[:, :HD], ctx.QKV[:, HD:2 * HD]
    Sbuf = torch.empty(ctx.n_blocks * NC * n_head * CP, CP, dtype=torch.float32, device=dev)
    blk_sz = NC * n_head * CP * CP
    for g in ctx.groups:
        A = L.mat(Qv[g.q0 * NP:], L.MAJOR_K, mn_off=(0, CP, g.q_si * NP, g.q_so * NP), k_off=(d, 0, 0, 0))
        B = L.mat(Kv[g.k0 * NP:], L.MAJOR_K, mn_off=(0, CP, g.k_si * NP, g.k_so * NP), k_off=(d, 0, 0, 0))
        D = L.out(Sbuf[g.blk0 * NC * n_head * CP:], CP, off=(CP * CP, n_head * CP * CP, blk_sz, g.n_in * blk_sz))
        L.gemm(A, B, D, CP, CP, d, nb=(n_head, NC, g.n_in, g.n_out), alpha=1.0 / math.sqrt(d))
    P = torch.empty(ctx.n_blocks * NC * n_head * CP, CP, dtype=dt, device=dev)
    rc = L.lib().csn_softmax_fwd(Sbuf.data_ptr(), P.data_ptr(), Sbuf.shape[0], CP, geom.kv_len, CP, geom.chunk,
                                 L.dtype_code(dt), L.stream_ptr())
    L.check(rc, "csn_softmax_fwd")
    return Sbuf, P


def use_fused_ln() -> bool:
    """CSN_FUSED_LN=0 selects the separate projection GEMM + csn_add_ln_fwd pair (tests cross-check the two)."""
    return os.environ.get("CSN_FUSED_LN", "1") != "0"


@dataclass(frozen=True)
class ChannelMajorResidual:
    """Where the fp32 residual of every slot lives in the reference's own (.., 256, N, 1) layout: slot s is the
    [256][N] matrix at bases[sel[s]] + off[s] elements, channel stride ch_stride (csa_models.py:92-94)."""
    bases: tuple          # up to two fp32 CUDA tensors (kept alive by the caller)
    sel: tuple            # per slot: index into bases
    off: tuple            # per slot: element offset
    ch_stride: int
    n_points: int


def attention_forward(Xh, Xf, groups, n_slots, n_blocks, w_q, w_k, w_v, w_o, gamma, beta, geom: Geometry,
                      n_head: int, want_colsum: bool = True, save_for_backward: bool = True,
                      want_y: bool = True, residual_cm: ChannelMajorResidual = None,
                      colsum_blocks: int = None, dropout_p: float = 0.0, seed: int = 0,
                      chunk_sum: torch.Tensor = None) -> AttnContext:
    """Forward of all blocks. Xh/Xf: packed slots [S*NP, 256] (16-bit / fp32). With residual_cm (and want_y False)
    Xf may be None: the output projection, the residual add and the LayerNorm statistics run as ONE kernel
    (csn_gemm_res_ln) that reads the residual from the channel-major inputs; colsum then covers the first
    colsum_blocks blocks only."""
    dev, dt = Xh.device, Xh.dtype
    NP, CP, NC = geom.rows_pad, geom.chunk_pad, geom.n_chunks
    HD = w_q.shape[0]
    d = HD // n_head
    ctx = AttnContext(geom, n_head, d, list(groups), n_slots, n_blocks, Xh=Xh, Xf=Xf)
    # training-mode dropout (csa_models.py:115,141): two independent masks (probabilities, fc output), regenerated by
    # the kernels from (seed, row, column); stored with the context so that the backward pass uses the same ones
    drop = (float(dropout_p), int(seed) & 0x7FFFFFFF, (int(seed) * 2654435761 + 97) & 0x7FFFFFFF) if dropout_p > 0.0 else (0.0, 0, 0)
    ctx.extra["drop"] = drop
    ctx.Wqkv16 = torch.cat([w_q, w_k, w_v], dim=0).to(dt).contiguous()  # [3HD, 256]
    ctx.Wo16 = w_o.to(dt).contiguous()                                   # [256, HD]
    ctx.gamma = gamma
    # --- projections: one GEMM for Q, K and V of every slot (csa_models.py:103-105, de-duplicated)
    QKV = torch.empty(n_slots * NP, 3 * HD, dtype=dt, device=dev)
    ragged = any(g.kv_lens or g.q_lens for g in groups)
    fused = use_fused_attention(d) and all(g.k0 == g.v0 and g.k_si == g.v_si and g.k_so == g.v_so for g in groups)
    fused_ln = residual_cm is not None and not want_y and use_fused_ln()
    # V centred on its per-chunk key mean c = mean_chunk(X) Wv^T (softmax rows sum to one: attn (V - c) + c == attn V):
    # the 16-bit V - c, the 16-bit O' = attn (V - c) and delta = rowsum(dO o O') lose ~10x less to rounding than V / O
    # (post-ReLU features give every value row a large common mean), so no rounding residual of O is kept; c Wo^T
    # re-enters the pre-LayerNorm sum in fp32 (zbias of csn_gemm_res_ln), d fc.weight gets the matching correction.
    center = fused and fused_ln and not ragged and chunk_sum is not None and use_centered_v() and drop[0] == 0.0
    if drop[0] > 0.0 and not fused:
        raise L.CsnError("training-mode dropout needs the fused attention kernels (d_head 64 or 256)")
    if center:
        cvec = torch.empty(n_slots * NC, HD, dtype=torch.float32, device=dev)
        sgemm_small(chunk_sum, w_v.detach().float().contiguous(), cvec, n_slots * NC, HD, 256, alpha=1.0 / geom.chunk)
        A_, B_, D_ = L.mat(Xh, L.MAJOR_K), L.mat(ctx.Wqkv16, L.MAJOR_K), L.out(QKV, 3 * HD)
        rc = L.lib().csn_gemm_colbias(C.byref(A_), C.byref(B_), C.byref(D_), n_slots * NP, 3 * HD, 256, 1.0, cvec.data_ptr(),
                                      HD, 2 * HD, CP, geom.chunk, L.stream_ptr())
        L.check(rc, "csn_gemm_colbias")
        kv_idx = cached_table(("kv_chunk", tuple(groups), n_blocks, NC), dev, lambda: _kv_chunk_table(groups, n_blocks, NC))
        zbias = torch.empty(n_blocks * NC, 256, dtype=torch.float32, device=dev)
        sgemm_small(cvec, w_o.detach().float().contiguous(), zbias, n_blocks * NC, 256, HD, a_rows=kv_idx)
        ctx.extra["center"] = (cvec, kv_idx)
    else:
        L.gemm(L.mat(Xh, L.MAJOR_K), L.mat(ctx.Wqkv16, L.MAJOR_K), L.out(QKV, 3 * HD), n_slots * NP, 3 * HD, 256)
        zbias = None
    ctx.QKV = QKV
    Qv, Kv, Vv = QKV[:, :HD], QKV[:, HD:2 * HD], QKV[:, 2 * HD:]
    O = (torch.zeros if ragged else torch.empty)(n_blocks * NP, HD, dtype=dt, device=dev)
    if fused:
        # --- fused flash-style core (csa_models.py:139-142): scores never leave TMEM
        items = attn_items(groups, geom, n_head, d, dev)
        lse = torch.empty(n_blocks * n_head * NP, dtype=torch.float32, device=dev)
        want_lo = (torch.is_grad_enabled() or save_for_backward) and not center and os.environ.get("CSN_O_LO", "1") == "1"
        O_lo = (torch.zeros_like(O) if ragged else torch.empty_like(O)) if want_lo else None
        rc = L.lib().csn_attn_fwd(Qv.data_ptr(), Kv.data_ptr(), Vv.data_ptr(), n_slots * NP, n_slots * NP, HD,
                                  3 * HD, 3 * HD, 3 * HD, d, L.dtype_code(dt), items.data_ptr(), items.shape[0],
                                  O.data_ptr(), O.shape[0], HD, lse.data_ptr(), O_lo.data_ptr() if O_lo is not None else None,
                                  _paired(geom), drop[1], drop[0], L.stream_ptr())
        L.check(rc, "csn_attn_fwd")
        ctx.extra["lse"] = lse
        ctx.extra["O_lo"] = O_lo
    elif ragged:
        raise L.CsnError("ragged blocks (per-block q_lens / kv_lens) need the fused attention kernels (d_head 64 or 256)")
    else:
        # --- materialised path: S = (Q K^T)/sqrt(d) -> softmax over the 500 valid keys -> O = P V
        Sbuf, P = _scores_and_probs(ctx_with(ctx, QKV))
        ctx.P = P
        ctx.extra["Sbuf"] = Sbuf  # scratch, reused for dP in backward
        prow = NC * n_head * CP  # rows of P per block
        for g in groups:
            A = L.mat(P[g.blk0 * prow:], L.MAJOR_K, mn_off=(CP, n_head * CP, prow, g.n_in * prow))
            B = L.mat(Vv[g.v0 * NP:], L.MAJOR_MN, mn_off=(d, 0, 0, 0), k_off=(0, CP, g.v_si * NP, g.v_so * NP))
            D = L.out(O[g.blk0 * NP:], HD, off=(d, CP * HD, NP * HD, g.n_in * NP * HD))
            L.gemm(A, B, D, CP, d, CP, nb=(n_head, NC, g.n_in, g.n_out))
    ctx.O = O
    # --- output projection (csa_models.py:115), residual + LayerNorm (:116-118), pooled column sums
    Z = torch.empty(n_blocks * NP, 256, dtype=torch.float32, device=dev)
    ctx.res_block = cached_table(("res_block", tuple(groups), n_blocks), dev,
                                 lambda: _res_block_table(groups, n_blocks))
    if fused_ln:
        r = residual_cm
        sel = cached_table(("res_sel", tuple(groups), n_blocks, r.sel), dev,
                           lambda: torch.tensor([r.sel[q] for q in _res_block_table(groups, n_blocks).tolist()], dtype=torch.int32))
        off = cached_table(("res_row", tuple(groups), n_blocks, r.off, r.ch_stride), dev,
                           lambda: torch.tensor([r.off[q] // r.ch_stride for q in _res_block_table(groups, n_blocks).tolist()], dtype=torch.int32))
        ctx.mean = torch.empty(n_blocks * NP, dtype=torch.float32, device=dev)
        ctx.rstd = torch.empty_like(ctx.mean)
        A, B = L.mat(O, L.MAJOR_K), L.mat(ctx.Wo16, L.MAJOR_K)
        b0, b1 = r.bases[0], r.bases[-1]
        rc = L.lib().csn_gemm_res_ln(C.byref(A), C.byref(B), Z.data_ptr(), 256, n_blocks * NP, HD, 1.0,
                                     b0.data_ptr(), b0.numel() // r.ch_stride, b1.data_ptr(), b1.numel() // r.ch_stride,
                                     sel.data_ptr(), off.data_ptr(), r.ch_stride, min(r.n_points, geom.n_points),
                                     NP, CP, geom.chunk, 1e-6, ctx.mean.data_ptr(), ctx.rstd.data_ptr(),
                                     zbias.data_ptr() if zbias is not None else None, drop[2], drop[0], L.stream_ptr())
        L.check(rc, "csn_gemm_res_ln")
        ctx.colsum = None
        if want_colsum:
            nb_cs = n_blocks if colsum_blocks is None else colsum_blocks
            parts = torch.empty(nb_cs * NP // 64, 256, dtype=torch.float32, device=dev)
            rc = L.lib().csn_ln_colsum(Z.data_ptr(), ctx.mean.data_ptr(), ctx.rstd.data_ptr(), gamma.data_ptr(),
                                       beta.data_ptr(), parts.data_ptr(), nb_cs * NP, NP, CP, geom.chunk, L.stream_ptr())
            L.check(rc, "csn_ln_colsum")
            ctx.colsum = torch.empty(nb_cs, 256, dtype=torch.float32, device=dev)
            rc = L.lib().csn_colsum_reduce(parts.data_ptr(), ctx.colsum.data_ptr(), nb_cs, NP // 64,
                                           1.0 / geom.n_points, L.stream_ptr())
            L.check(rc, "csn_colsum_reduce")
        ctx.Z, ctx.Y = Z, None
        ctx.beta = beta
        return ctx
    if Xf is None:
        raise L.CsnError("attention_forward: the fp32 row copy Xf is required when the fused residual/LayerNorm epilogue is off")
    L.gemm(L.mat(O, L.MAJOR_K), L.mat(ctx.Wo16, L.MAJOR_K), L.out(Z, 256), n_blocks * NP, 256, HD)
    Y = torch.empty_like(Z) if want_y else None   # CSA: consumers re-normalise z on the fly, Y never hits HBM
    ctx.mean = torch.empty(n_blocks * NP, dtype=torch.float32, device=dev)
    ctx.rstd = torch.empty_like(ctx.mean)
    parts = torch.empty(n_blocks * NP // 64, 256, dtype=torch.float32, device=dev) if want_colsum else None
    rc = L.lib().csn_add_ln_fwd(Z.data_ptr(), Xf.data_ptr(), ctx.res_block.data_ptr(),
                                Y.data_ptr() if Y is not None else None, None,
                                ctx.mean.data_ptr(), ctx.rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                parts.data_ptr() if want_colsum else None, n_blocks * NP, NP, CP, geom.chunk,
                                1e-6, L.dtype_code(dt), drop[2], drop[0], L.stream_ptr())
    L.check(rc, "csn_add_ln_fwd")
    ctx.colsum = None
    if want_colsum:   # pooled mean over the points of every block (csa_models.py:212,219), fixed summation order
        ctx.colsum = torch.empty(n_blocks, 256, dtype=torch.float32, device=dev)
        rc = L.lib().csn_colsum_reduce(parts.data_ptr(), ctx.colsum.data_ptr(), n_blocks, NP // 64,
                                       1.0 / geom.n_points, L.stream_ptr())
        L.check(rc, "csn_colsum_reduce")
    ctx.Z, ctx.Y = Z, Y
    ctx.beta = beta
    return ctx


def _attention_backward_materialised(ctx: AttnContext, dO: torch.Tensor, dQKV: torch.Tensor) -> None:
    """dQ|dK|dV through materialised P / dP / dS tiles (csn_gemm + csn_softmax_bwd): the fallback for
    head sizes the fused kernels do not cover, and the cross-check of the fused path in the tests."""
    geom, h, d = ctx.geom, ctx.n_head, ctx.d_head
    NP, CP, NC = geom.rows_pad, geom.chunk_pad, geom.n_chunks
    HD = h * d
    dt = ctx.Xh.dtype
    lib = L.lib()
    Qv, Kv, Vv = ctx.QKV[:, :HD], ctx.QKV[:, HD:2 * HD], ctx.QKV[:, 2 * HD:]
    dQv, dKv, dVv = dQKV[:, :HD], dQKV[:, HD:2 * HD], dQKV[:, 2 * HD:]
    if ctx.P is None:   # fused forward: rebuild the probabilities
        ctx.extra["Sbuf"], ctx.P = _scores_and_probs(ctx)
    # --- dP = dO V^T per (block, chunk, head)
    dP = ctx.extra["Sbuf"]
    blk_sz = NC * h * CP * CP
    prow = NC * h * CP
    for g in ctx.groups:
        A = L.mat(dO[g.blk0 * NP:], L.MAJOR_K, mn_off=(0, CP, NP, g.n_in * NP), k_off=(d, 0, 0, 0))
        B = L.mat(Vv[g.v0 * NP:], L.MAJOR_K, mn_off=(0, CP, g.v_si * NP, g.v_so * NP), k_off=(d, 0, 0, 0))
        D = L.out(dP[g.blk0 * prow:], CP, off=(CP * CP, h * CP * CP, blk_sz, g.n_in * blk_sz))
        L.gemm(A, B, D, CP, CP, d, nb=(h, NC, g.n_in, g.n_out))
    # --- dS = P o (dP - rowsum(P o dP)) / sqrt(d)
    dS = torch.empty_like(ctx.P)
    rc = lib.csn_softmax_bwd(ctx.P.data_ptr(), dP.data_ptr(), dS.data_ptr(), dP.shape[0], CP, geom.kv_len, CP,
                             geom.chunk, 1.0 / math.sqrt(d), L.dtype_code(dt), L.stream_ptr())
    L.check(rc, "csn_softmax_bwd")
    for g in ctx.groups:
        nb = (h, NC, g.n_in, g.n_out)
        Pk = L.mat(ctx.P[g.blk0 * prow:], L.MAJOR_MN, k_off=(CP, h * CP, prow, g.n_in * prow))        # P^T
        dSk = L.mat(dS[g.blk0 * prow:], L.MAJOR_K, mn_off=(CP, h * CP, prow, g.n_in * prow))
        dSt = L.mat(dS[g.blk0 * prow:], L.MAJOR_MN, k_off=(CP, h * CP, prow, g.n_in * prow))          # dS^T
        dOm = L.mat(dO[g.blk0 * NP:], L.MAJOR_MN, mn_off=(d, 0, 0, 0), k_off=(0, CP, NP, g.n_in * NP))
        Km = L.mat(Kv[g.k0 * NP:], L.MAJOR_MN, mn_off=(d, 0, 0, 0), k_off=(0, CP, g.k_si * NP, g.k_so * NP))
        Qm = L.mat(Qv[g.q0 * NP:], L.MAJOR_MN, mn_off=(d, 0, 0, 0), k_off=(0, CP, g.q_si * NP, g.q_so * NP))
        off = (d, CP * 3 * HD, NP * 3 * HD, g.n_in * NP * 3 * HD)
        L.gemm(Pk, dOm, L.out(dVv[g.blk0 * NP:], 3 * HD, off=off), CP, d, CP, nb=nb)    # dV = P^T dO
        L.gemm(dSk, Km, L.out(dQv[g.blk0 * NP:], 3 * HD, off=off), CP, d, CP, nb=nb)    # dQ = dS K
        L.gemm(dSt, Qm, L.out(dKv[g.blk0 * NP:], 3 * HD, off=off), CP, d, CP, nb=nb)    # dK = dS^T Q


def _pick_split(tiles: int, kb_total: int, target: int = 296) -> int:
    """split-K factor such that tiles*split is about two waves of 148 CTAs, each split >= 8 k-blocks."""
    return max(1, min(-(-target // max(tiles, 1)), max(1, kb_total // 8)))


def _attention_backward_ds(ctx, lib, dO, dQKV, lse, delta, drop, ragged):
    """dV kernel + dS to HBM (+ dQ inside the kernel at d_head 64) + dK = dS^T Q (and dQ = dS K) as GEMMs: the
    d_head-256 path (S, dP and a [128 x 256] dQ accumulator do not fit in TMEM together)."""
    geom, h, d = ctx.geom, ctx.n_head, ctx.d_head
    NP, CP, NC = geom.rows_pad, geom.chunk_pad, geom.n_chunks
    HD = h * d
    S, nblk = ctx.n_slots, ctx.n_blocks
    dev, dt = ctx.Xh.device, ctx.Xh.dtype
    Qv, Kv, Vv = ctx.QKV[:, :HD], ctx.QKV[:, HD:2 * HD], ctx.QKV[:, 2 * HD:]
    dQv, dKv, dVv = dQKV[:, :HD], dQKV[:, HD:2 * HD], dQKV[:, 2 * HD:]
    prow = NC * h * CP
    it_dv = attn_items(ctx.groups, geom, h, d, dev, "dv")
    rc = lib.csn_attn_bwd_dv(Kv.data_ptr(), Qv.data_ptr(), dO.data_ptr(), S * NP, S * NP, nblk * NP, HD, 3 * HD, 3 * HD, HD,
                             d, L.dtype_code(dt), it_dv.data_ptr(), it_dv.shape[0], dVv.data_ptr(), nblk * NP, 3 * HD,
                             lse.data_ptr(), _paired(geom), drop[1], drop[0], L.stream_ptr())
    L.check(rc, "csn_attn_bwd_dv")
    # the dQ kernel writes the key tiles it visits; columns beyond them must read as zeros in dS^T Q
    if ragged:
        dS = _zeroed_workspace(nblk * prow * CP, dt, dev).view(nblk * prow, CP)
    else:
        alloc = torch.empty if ((geom.kv_len + 127) // 128) * 128 >= CP else torch.zeros
        dS = alloc(nblk * prow, CP, dtype=dt, device=dev)
    it_dq = attn_items(ctx.groups, geom, h, d, dev, "dq")
    # dQ inside the kernel: slower at d_head 256 (re-reads K_j through a shallow ring), faster at d_head 64
    fuse_dq = os.environ.get("CSN_FUSED_DQ", "1" if d == 64 else "0") == "1"
    rc = lib.csn_attn_bwd_dq(Qv.data_ptr(), dO.data_ptr(), Kv.data_ptr(), Vv.data_ptr(), S * NP, nblk * NP, S * NP,
                             HD, 3 * HD, HD, 3 * HD, 3 * HD, d, L.dtype_code(dt), it_dq.data_ptr(),
                             it_dq.shape[0], dQv.data_ptr() if fuse_dq else None, 3 * HD, dS.data_ptr(),
                             dS.shape[0], CP, lse.data_ptr(), delta.data_ptr(), _paired(geom), drop[1], drop[0],
                             L.stream_ptr())
    L.check(rc, "csn_attn_bwd_dq")
    for g in ctx.groups:
        nb = (h, NC, g.n_in, g.n_out)
        dSk = L.mat(dS[g.blk0 * prow:], L.MAJOR_K, mn_off=(CP, h * CP, prow, g.n_in * prow))
        dSt = L.mat(dS[g.blk0 * prow:], L.MAJOR_MN, k_off=(CP, h * CP, prow, g.n_in * prow))          # dS^T
        Km = L.mat(Kv[g.k0 * NP:], L.MAJOR_MN, mn_off=(d, 0, 0, 0), k_off=(0, CP, g.k_si * NP, g.k_so * NP))
        Qm = L.mat(Qv[g.q0 * NP:], L.MAJOR_MN, mn_off=(d, 0, 0, 0), k_off=(0, CP, g.q_si * NP, g.q_so * NP))
        off = (d, CP * 3 * HD, NP * 3 * HD, g.n_in * NP * 3 * HD)
        if not fuse_dq and d > 128 and os.environ.get("CSN_GEMM_DUAL", "1") != "0":
            # dQ = dS K and dK = dS^T Q in ONE launch, tiles of the same (block, chunk) back to back: the second
            # read of the dS tile comes from L2 (csn_gemm_dual)
            nb4 = (C.c_int32 * 4)(*[int(v) for v in nb])
            D0 = L.out(dQv[g.blk0 * NP:], 3 * HD, off=off)
            D1 = L.out(dKv[g.blk0 * NP:], 3 * HD, off=off)
            rc = lib.csn_gemm_dual(C.byref(dSk), C.byref(Km), C.byref(D0), C.byref(dSt), C.byref(Qm), C.byref(D1),
                                   CP, d, CP, nb4, 1.0, L.stream_ptr())
            L.check(rc, "csn_gemm_dual")
            continue
        if not fuse_dq:
            L.gemm(dSk, Km, L.out(dQv[g.blk0 * NP:], 3 * HD, off=off), CP, d, CP, nb=nb)    # dQ = dS K
        L.gemm(dSt, Qm, L.out(dKv[g.blk0 * NP:], 3 * HD, off=off), CP, d, CP, nb=nb)    # dK = dS^T Q


_WORKSPACES: dict = {}


def _zeroed_workspace(numel: int, dtype, device) -> torch.Tensor:
    """A persistent buffer that was zero when first handed out and is NOT cleared again.  For the ragged dS buffer
    (gigabytes at MinkowskiNet sizes: clearing it cost 1.9 ms of a 13 ms step): the dQ kernel writes the tiles it
    visits; everything else is multiplied by exactly-zero padded Q rows in dS^T Q or lands in padded dK rows that
    nobody reads, so any FINITE stale value there is harmless -- only uninitialised NaN/Inf patterns would not be.
    `reset_workspaces()` drops the buffers (e.g. after a step that overflowed)."""
    key = (str(device), dtype)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < numel:
        buf = torch.zeros(numel, dtype=dtype, device=device)
        _WORKSPACES[key] = buf
    return buf[:numel]


def reset_workspaces() -> None:
    _WORKSPACES.clear()


def attention_backward(ctx: AttnContext, dY: torch.Tensor, need_dx: bool, amax: torch.Tensor = None,
                       bcast: torch.Tensor = None, bcast_idx: torch.Tensor = None, bcast_scale: float = 0.0,
                       src_idx: torch.Tensor = None, src_w: torch.Tensor = None, split_v: bool = False):
    """Backward of attention_forward. dY: [blocks*NP, 256] fp32 (zero in pad rows).
    Returns dict with dWq, dWk, dWv, dWo (fp32, reference layouts), dgamma, dbeta and, if need_dx,
    dX [S*NP, 256] fp32 (padded row-major, gradient w.r.t. every slot's features)."""
    geom, h, d = ctx.geom, ctx.n_head, ctx.d_head
    NP, CP, NC = geom.rows_pad, geom.chunk_pad, geom.n_chunks
    HD = h * d
    dev, dt = ctx.Xh.device, ctx.Xh.dtype
    nblk, S = ctx.n_blocks, ctx.n_slots
    lib = L.lib()
    # --- dynamic range: the backward pass is linear in dY, so it is run on s*dY with s a power of two
    #     chosen on the device (no host sync) such that max|s*dY| = 64..128; 16-bit intermediates
    #     (dZ, dO, dS, dQ|dK|dV) then sit in the normal range of fp16 instead of its subnormals.
    #     amax comes from the kernel that produced dY (csn_combine_bwd) or is reduced here; the scale
    #     itself is applied inside csn_ln_bwd when dY is loaded.
    if amax is None:
        amax = dY.abs().max().reshape(1)
    # every parameter gradient of the step lives in ONE zero-initialised flat buffer (one fill, one csn_grad_unscale):
    # [dgamma 256 | dbeta 256 | dWo 256*HD | dWqkv 3*HD*256]
    gflat = torch.zeros(512 + 4 * HD * 256, dtype=torch.float32, device=dev)
    # --- LayerNorm backward
    dZ = torch.empty_like(ctx.Z) if need_dx else None
    dZ16 = torch.empty(nblk * NP, 256, dtype=dt, device=dev)
    dgamma, dbeta = gflat[:256], gflat[256:512]
    center = ctx.extra.get("center")
    drop = ctx.extra.get("drop", (0.0, 0, 0))
    gsum = torch.zeros(nblk * NC, 256, dtype=torch.float32, device=dev) if center is not None else None
    rc = lib.csn_ln_bwd(dY.data_ptr(), ctx.Z.data_ptr(), ctx.mean.data_ptr(), ctx.rstd.data_ptr(),
                        ctx.gamma.data_ptr(), dZ.data_ptr() if dZ is not None else None, dZ16.data_ptr(),
                        dgamma.data_ptr(), dbeta.data_ptr(), nblk * NP, NP, CP, geom.chunk, L.dtype_code(dt),
                        amax.data_ptr(), bcast.data_ptr() if bcast is not None else None,
                        bcast_idx.data_ptr() if bcast_idx is not None else None, bcast_scale,
                        src_idx.data_ptr() if src_idx is not None else None,
                        src_w.data_ptr() if src_w is not None else None,
                        gsum.data_ptr() if gsum is not None else None, drop[2], drop[0], L.stream_ptr())
    L.check(rc, "csn_ln_bwd")
    split = _pick_split(2 * ((HD + 255) // 256), nblk * NP // 64)
    # --- dWo = dZ^T O  (contraction over all rows; both operands consumed MN-major)
    dWo = gflat[512:512 + 256 * HD].view(256, HD)
    L.gemm(L.mat(dZ16, L.MAJOR_MN), L.mat(ctx.O, L.MAJOR_MN), L.out(dWo, HD, accumulate=True), 256, HD, nblk * NP,
           split_k=split)
    if center is not None:   # ctx.O holds O' = O - c: d fc.weight += sum over chunks of (sum_rows dZ)^T c
        cvec, kv_idx = center
        sgemm_small(gsum, cvec, dWo, 256, HD, nblk * NC, b_rows=kv_idx, trans_a=True, trans_b=True, accumulate=True)
    # --- dO = dZ Wo; on the fused path the same kernel also forms delta = rowsum(dO o O) per (row, head)
    dO = torch.empty(nblk * NP, HD, dtype=dt, device=dev)
    fused_delta = (ctx.P is None and "lse" in ctx.extra and os.environ.get("CSN_FUSED_BWD", "1") != "0"
                   and HD > 128 and os.environ.get("CSN_FUSED_DELTA", "1") != "0")
    delta = None
    if fused_delta:
        delta = torch.empty_like(ctx.extra["lse"])
        O_lo = ctx.extra.get("O_lo")
        Am, Bm = L.mat(dZ16, L.MAJOR_K), L.mat(ctx.Wo16, L.MAJOR_MN)
        rc = lib.csn_gemm_delta(C.byref(Am), C.byref(Bm), dO.data_ptr(), HD, nblk * NP, HD, 256, 1.0, ctx.O.data_ptr(),
                                O_lo.data_ptr() if O_lo is not None else None, HD, delta.data_ptr(), NP, h, d,
                                L.stream_ptr())
        L.check(rc, "csn_gemm_delta")
    else:
        L.gemm(L.mat(dZ16, L.MAJOR_K), L.mat(ctx.Wo16, L.MAJOR_MN), L.out(dO, HD), nblk * NP, HD, 256)
    Qv, Kv, Vv = ctx.QKV[:, :HD], ctx.QKV[:, HD:2 * HD], ctx.QKV[:, 2 * HD:]
    fused_bwd = ctx.P is None and "lse" in ctx.extra and os.environ.get("CSN_FUSED_BWD", "1") != "0"
    ragged = any(g.kv_lens or g.q_lens for g in ctx.groups)
    dQKV = (torch.zeros if ragged else torch.empty)(nblk * NP, 3 * HD, dtype=dt, device=dev)
    dQv, dKv, dVv = dQKV[:, :HD], dQKV[:, HD:2 * HD], dQKV[:, 2 * HD:]
    prow = NC * h * CP
    if fused_bwd:
        # --- fused attention backward: delta, dV (P^T rebuilt from lse), dQ (+ dS to HBM), dK = dS^T Q
        lse = ctx.extra["lse"]
        if delta is None:
            delta = torch.empty_like(lse)
            O_lo = ctx.extra.get("O_lo")
            rc = lib.csn_attn_delta(dO.data_ptr(), ctx.O.data_ptr(), O_lo.data_ptr() if O_lo is not None else None,
                                    delta.data_ptr(), nblk * NP, NP, h, d, HD, L.dtype_code(dt), L.stream_ptr())
            L.check(rc, "csn_attn_delta")
        # d_head 64: dK and dV from ONE key-stationary kernel, dQ from the query-stationary one, dS never materialised
        # (S^T, dP^T and both output accumulators fit in TMEM: attn_dkv.cu).  CSN_FUSED_DKV=0 keeps the older
        # dV kernel + dS buffer + dK GEMM for cross-checks.
        fuse_dkv = d == 64 and os.environ.get("CSN_FUSED_DKV", "1") == "1" and os.environ.get("CSN_FUSED_DQ", "1") == "1"
        if fuse_dkv:
            it_kv = attn_items(ctx.groups, geom, h, d, dev, "dkv")
            stat_scratch = torch.empty(2 * lse.numel(), dtype=torch.float32, device=dev)
            rc = lib.csn_attn_bwd_dkv(Kv.data_ptr(), Vv.data_ptr(), Qv.data_ptr(), dO.data_ptr(), S * NP, S * NP, nblk * NP, HD,
                                      3 * HD, 3 * HD, 3 * HD, HD, d, L.dtype_code(dt), it_kv.data_ptr(), it_kv.shape[0],
                                      dKv.data_ptr(), dVv.data_ptr(), nblk * NP, 3 * HD, lse.data_ptr(), delta.data_ptr(),
                                      lse.numel(), stat_scratch.data_ptr(), drop[1], drop[0], L.stream_ptr())
            L.check(rc, "csn_attn_bwd_dkv")
            it_dq = attn_items(ctx.groups, geom, h, d, dev, "dq")
            rc = lib.csn_attn_bwd_dq(Qv.data_ptr(), dO.data_ptr(), Kv.data_ptr(), Vv.data_ptr(), S * NP, nblk * NP, S * NP,
                                     HD, 3 * HD, HD, 3 * HD, 3 * HD, d, L.dtype_code(dt), it_dq.data_ptr(),
                                     it_dq.shape[0], dQv.data_ptr(), 3 * HD, None, 0, CP, lse.data_ptr(), delta.data_ptr(),
                                     _paired(geom), drop[1], drop[0], L.stream_ptr())
            L.check(rc, "csn_attn_bwd_dq")
        else:
            _attention_backward_ds(ctx, lib, dO, dQKV, lse, delta, drop, ragged)
    else:
        if drop[0] > 0.0:
            raise L.CsnError("training-mode dropout needs the fused attention backward (CSN_FUSED_BWD=1)")
        _attention_backward_materialised(ctx, dO, dQKV)
    # --- projection weight gradients: dW = sum_blocks dProj^T X[slot]  (contraction over points, both
    #     operands consumed MN-major; split-K sized for ~2 waves of CTAs, fp32 atomics into dWqkv)
    dWqkv = gflat[512 + 256 * HD:].view(3 * HD, 256)
    Xh = ctx.Xh
    for g in ctx.groups:
        same = (g.q0, g.q_si, g.q_so) == (g.k0, g.k_si, g.k_so) == (g.v0, g.v_si, g.v_so)
        if same and g.n_out == 1 and g.q_si == 1:
            # self blocks over consecutive slots: one long contraction for dWq|dWk|dWv together
            Kdim = g.n_in * NP
            A = L.mat(dQKV[g.blk0 * NP:(g.blk0 + g.n_in) * NP], L.MAJOR_MN)
            B = L.mat(Xh[g.q0 * NP:(g.q0 + g.n_in) * NP], L.MAJOR_MN)
            L.gemm(A, B, L.out(dWqkv, 256, accumulate=True), 3 * HD, 256, Kdim,
                   split_k=_pick_split((3 * HD + 127) // 128, Kdim // 64))
            continue
        nb = (g.n_in, g.n_out, 1, 1)
        nbt = g.n_in * g.n_out
        roles = [(dQv, dWqkv[:HD], HD, g.q0, g.q_si, g.q_so)]
        if (g.k0, g.k_si, g.k_so) == (g.v0, g.v_si, g.v_so):
            roles.append((dQKV[:, HD:], dWqkv[HD:], 2 * HD, g.k0, g.k_si, g.k_so))
        else:
            roles.append((dKv, dWqkv[HD:2 * HD], HD, g.k0, g.k_si, g.k_so))
            roles.append((dVv, dWqkv[2 * HD:], HD, g.v0, g.v_si, g.v_so))
        for (dproj, dst, Mdim, s0, si, so) in roles:
            A = L.mat(dproj[g.blk0 * NP:], L.MAJOR_MN, k_off=(NP, g.n_in * NP))
            B = L.mat(Xh[s0 * NP:], L.MAJOR_MN, k_off=(si * NP, so * NP))
            L.gemm(A, B, L.out(dst, 256, accumulate=True), Mdim, 256, NP, nb=nb,
                   split_k=_pick_split(nbt * ((Mdim + 127) // 128), NP // 64))
    grads = {"dWq": dWqkv[:HD], "dWk": dWqkv[HD:2 * HD], "dWv": dWqkv[2 * HD:], "dWo": dWo,
             "dgamma": dgamma, "dbeta": dbeta}
    if need_dx:
        dX = torch.zeros(S * NP, 256, dtype=torch.float32, device=dev)
        # split_v: the gradient that reaches a slot through its VALUE role goes to a separate buffer (callers whose
        # key and value inputs are distinct autograd tensors holding the same data)
        dXv = torch.zeros_like(dX) if split_v else dX
        Wq16, Wk16, Wv16 = ctx.Wqkv16[:HD], ctx.Wqkv16[HD:2 * HD], ctx.Wqkv16[2 * HD:]
        for g in ctx.groups:
            nb = (g.n_in, g.n_out, 1, 1)
            q_sl, k_sl, v_sl = (g.q0, g.q_si, g.q_so), (g.k0, g.k_si, g.k_so), (g.v0, g.v_si, g.v_so)
            # roles that land in the same slots are ONE contraction over their concatenated columns of dQKV / rows of
            # Wqkv (self blocks: K = 3 HD; cross blocks: dQ alone, dK|dV together) instead of one launch per role
            if not split_v and q_sl == k_sl == v_sl:
                roles = ((dQKV, ctx.Wqkv16, 3 * HD, q_sl, dX),)
            elif not split_v and k_sl == v_sl:
                roles = ((dQv, Wq16, HD, q_sl, dX), (dQKV[:, HD:], ctx.Wqkv16[HD:], 2 * HD, k_sl, dX))
            else:
                roles = ((dQv, Wq16, HD, q_sl, dX), (dKv, Wk16, HD, k_sl, dX), (dVv, Wv16, HD, v_sl, dXv))
            for (dproj, W16, Kdim, (s0, si, so), dst) in roles:
                A = L.mat(dproj[g.blk0 * NP:], L.MAJOR_K, mn_off=(NP, g.n_in * NP))
                B = L.mat(W16, L.MAJOR_MN)
                D = L.out(dst[s0 * NP:], 256, off=(si * NP * 256, so * NP * 256), accumulate=True)
                L.gemm(A, B, D, NP, 256, Kdim, nb=nb)
        if split_v:
            grads["dXv"] = dXv
        # residual path: dX[q slot] += dZ[block], and the loss scale taken out in the same pass
        rc = lib.csn_block_add(dZ.data_ptr(), ctx.res_block.data_ptr(), nblk, dX.data_ptr(), S, NP * 256, amax.data_ptr(),
                               L.stream_ptr())
        L.check(rc, "csn_block_add")
        grads["dX"] = dX
    rc = lib.csn_grad_unscale(gflat.data_ptr(), gflat.numel(), amax.data_ptr(), L.stream_ptr())
    L.check(rc, "csn_grad_unscale")
    if "dXv" in grads:
        rc = lib.csn_grad_unscale(grads["dXv"].data_ptr(), grads["dXv"].numel(), amax.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_grad_unscale")
    return grads
