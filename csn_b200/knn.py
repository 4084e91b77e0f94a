"""Host side of the shape-compatibility retrieval path (kNN graph over a shape collection).

Mirrors CrossShapeAt.get_retrieval_measure / get_knn_graph (MID-FC/csa_models.py:244-280), the
big-class variant (:360-404, candidates = a subset of the collection) and
HRNetSimCSN.cosine_similarity (MinkowskiNet/models/hrnet.py:472-490).  All arithmetic runs in
csrc/knn.cu through the C ABI; this file only builds the small work tables.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import _lib as L

TILE_ROWS = 128
NUM_SMS = 148
TOPK_LIMIT = 64   # csn_topk_rows: per-lane candidate lists of 8 / 16 / 32 / 64 entries


@dataclass
class ShapeStore:
    """Unit-norm 16-bit point features of a set of shapes, packed row-wise: shape s occupies rows
    [row0[s], row0[s] + length[s]) of `rows` (total_rows, 256).  This is the GPU-resident candidate
    store of SURVEY.md §8e (41 GB fp32 -> 20 GB 16-bit at S = 4000, N = 10k)."""
    rows: torch.Tensor
    row0: list[int]
    length: list[int]
    rows_lo: torch.Tensor = None   # optional fp16 residual (x - fp16(x)) * 2^11 for the exact re-score

    @property
    def n_shapes(self) -> int:
        return len(self.row0)

    def subset(self, idx) -> "ShapeStore":
        idx = [int(i) for i in idx]
        return ShapeStore(self.rows, [self.row0[i] for i in idx], [self.length[i] for i in idx], self.rows_lo)

    def tables(self):
        """Device copies of (row0, length, item0) as int32 tensors, built once per store: item0[s] = index of the first
        128-row tile of shape s in the store's tile enumeration (exclusive prefix sum of the tile counts)."""
        t = getattr(self, "_tables", None)
        if t is None:
            import numpy as np
            row0 = np.asarray(self.row0, dtype=np.int64)
            length = np.asarray(self.length, dtype=np.int64)
            nt = (length + TILE_ROWS - 1) // TILE_ROWS
            item0 = np.concatenate([[0], np.cumsum(nt)[:-1]]) if len(nt) else np.zeros(0, dtype=np.int64)
            dev = self.rows.device
            # one pinned staging buffer, asynchronous upload: building the tables never synchronises with the stream
            host = torch.from_numpy(np.stack([row0, length, item0]).astype(np.int32))
            if dev.type == "cuda":
                host = host.pin_memory()
            d = host.to(dev, non_blocking=True)
            t = (d[0], d[1], d[2], int(nt.sum()), host)
            object.__setattr__(self, "_tables", t)
        return t


def _normalize_into(src: torch.Tensor, dst: torch.Tensor, eps: float) -> None:
    assert src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()
    rows, D = src.shape
    rc = L.lib().csn_normalize_rows(src.data_ptr(), dst.data_ptr(), rows, D, eps, L.dtype_code(dst.dtype),
                                    L.stream_ptr())
    L.check(rc, "csn_normalize_rows")


def _normalize_split_into(src: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor, eps: float) -> None:
    rows, D = src.shape
    rc = L.lib().csn_normalize_rows_split(src.data_ptr(), hi.data_ptr(), lo.data_ptr(), rows, D, eps, L.stream_ptr())
    L.check(rc, "csn_normalize_rows_split")


def build_store(feats, dtype: torch.dtype = torch.float16, eps: float = 1e-12, exact: bool = False) -> ShapeStore:
    """feats: (S, N, D) fp32 CUDA tensor, or a list of (L_s, D) fp32 CUDA tensors (ragged).
    exact=True (fp16 only) also keeps the rounding residual of every row, which the exact re-scoring of
    the top-K boundary band needs."""
    if isinstance(feats, torch.Tensor):
        assert feats.dim() == 3
        S, N, D = feats.shape
        src = feats.contiguous().view(S * N, D)
        dst = torch.empty(S * N, D, dtype=dtype, device=feats.device)
        lo = None
        if exact and dtype == torch.float16:
            lo = torch.empty_like(dst)
            _normalize_split_into(src, dst, lo, eps)
        else:
            _normalize_into(src, dst, eps)
        return ShapeStore(dst, [s * N for s in range(S)], [N] * S, lo)
    lens = [int(f.shape[0]) for f in feats]
    D = int(feats[0].shape[1])
    total = sum(lens)
    dst = torch.empty(total, D, dtype=dtype, device=feats[0].device)
    row0, r = [], 0
    for f, n in zip(feats, lens):
        _normalize_into(f.contiguous(), dst[r:r + n], eps)
        row0.append(r)
        r += n
    return ShapeStore(dst, row0, lens)


def concat_stores(stores) -> ShapeStore:
    """One store from several (e.g. built batch by batch while the collection streams through the SSA layer)."""
    stores = list(stores)
    rows = torch.cat([s.rows for s in stores], dim=0)
    lo = torch.cat([s.rows_lo for s in stores], dim=0) if all(s.rows_lo is not None for s in stores) else None
    row0, length, base = [], [], 0
    for s in stores:
        row0 += [base + r for r in s.row0]
        length += list(s.length)
        base += s.rows.shape[0]
    return ShapeStore(rows, row0, length, lo)


def _split_for_balance(n_items: int, n_cand: int) -> int:
    """Number of candidate-list segments per (query, tile) so that the persistent grid's last wave
    is nearly full."""
    best, best_waste = 1, 1.0
    for ns in range(1, min(16, n_cand) + 1):
        tot = n_items * ns
        waste = (math.ceil(tot / NUM_SMS) * NUM_SMS - tot) / tot
        if waste < best_waste - 1e-9:
            best, best_waste = ns, waste
        if best_waste < 0.03:
            break
    return best


def scores_from_stores(q: ShapeStore, c: ShapeStore, query_block: int = 148) -> torch.Tensor:
    """(Sq, Sc) fp32 retrieval measure of every query shape against every candidate shape."""
    dev = q.rows.device
    Sq, Sc = q.n_shapes, c.n_shapes
    scores = torch.empty(Sq, Sc, dtype=torch.float32, device=dev)
    import numpy as np
    c_row0, c_len = c.tables()[:2]
    q.tables()   # (uploaded here so that refine_band finds them cached)
    cands = torch.stack([c_row0, c_len], dim=1).contiguous()
    lib = L.lib()
    q_row0_np, q_len_np = np.asarray(q.row0, dtype=np.int64), np.asarray(q.length, dtype=np.int64)
    for q0 in range(0, Sq, query_block):
        q1 = min(Sq, q0 + query_block)
        # tables (numpy, no Python loops over tiles): items ordered (query, tile, segment); partial rows (query, tile)
        lens = q_len_np[q0:q1]
        nts = (lens + TILE_ROWS - 1) // TILE_ROWS
        prow0 = np.concatenate([[0], np.cumsum(nts)[:-1]])
        part_rows = int(nts.sum())
        row_of_query = list(zip(prow0.tolist(), nts.tolist()))
        ns = _split_for_balance(part_rows, Sc)
        seg_b = np.arange(ns, dtype=np.int64) * Sc // ns
        seg_e = (np.arange(ns, dtype=np.int64) + 1) * Sc // ns
        qi = np.repeat(np.arange(q1 - q0), nts)                                  # query of every (query, tile) row
        ti = np.arange(part_rows) - np.repeat(prow0, nts)                       # tile index inside its query
        nvalid = np.minimum(TILE_ROWS, lens[qi] - ti * TILE_ROWS)
        tab = np.empty((part_rows, ns, 6), dtype=np.int32)
        tab[:, :, 0] = (q_row0_np[q0:q1][qi] + ti * TILE_ROWS)[:, None]
        tab[:, :, 1] = nvalid[:, None]
        tab[:, :, 2] = seg_b[None, :]
        tab[:, :, 3] = (seg_e - seg_b)[None, :]
        tab[:, :, 4] = np.arange(part_rows)[:, None] * Sc + seg_b[None, :]
        tab[:, :, 5] = 0
        items = tab.reshape(-1, 6)
        items_t = torch.from_numpy(items).pin_memory().to(dev, non_blocking=True)
        partial = torch.empty(part_rows * Sc, dtype=torch.float32, device=dev)
        rc = lib.csn_knn_scores(q.rows.data_ptr(), q.rows.shape[0], c.rows.data_ptr(), c.rows.shape[0],
                                L.dtype_code(q.rows.dtype), items_t.data_ptr(), len(items), cands.data_ptr(),
                                partial.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_knn_scores")
        uniform = len({nt for _, nt in row_of_query}) == 1 and len({q.length[s] for s in range(q0, q1)}) == 1
        if uniform:
            nt = row_of_query[0][1]
            rc = lib.csn_knn_reduce(partial.data_ptr(), scores[q0:q1].data_ptr(), q1 - q0, Sc, nt,
                                    q.length[q0], scores.stride(0), L.stream_ptr())
            L.check(rc, "csn_knn_reduce")
        else:
            for s, (prow, nt) in zip(range(q0, q1), row_of_query):
                rc = lib.csn_knn_reduce(partial[prow * Sc:].data_ptr(), scores[s:s + 1].data_ptr(), 1, Sc, nt,
                                        q.length[s], scores.stride(0), L.stream_ptr())
                L.check(rc, "csn_knn_reduce")
        # keep the tables alive until the kernels that read them have been enqueued on this stream
        del items_t, partial
    return scores


def topk_rows(scores: torch.Tensor, k: int):
    """Values and int64 indices of the k largest entries per row, sorted descending."""
    assert scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2 and scores.stride(1) == 1
    n_rows, n_cols = scores.shape
    val = torch.empty(n_rows, k, dtype=torch.float32, device=scores.device)
    idx = torch.empty(n_rows, k, dtype=torch.int64, device=scores.device)
    rc = L.lib().csn_topk_rows(scores.data_ptr(), scores.stride(0), n_rows, n_cols, k, val.data_ptr(),
                               idx.data_ptr(), L.stream_ptr())
    L.check(rc, "csn_topk_rows")
    return val, idx


def refine_band(scores: torch.Tensor, q: ShapeStore, c: ShapeStore, k: int, margin: float = 4e-5,
                query_block: int = 64) -> torch.Tensor:
    """Exact re-score of the top-k boundary band, entirely on the device.  For every query the candidates whose
    coarse score is within `margin` of (or above) the k-th best are selected by csn_knn_band_select — which writes the
    work tables of csn_knn_scores_exact itself — re-scored with split fp16 operands (22-bit significands) and patched
    into `scores` in place by csn_knn_band_patch: the top-k index sets then match an fp32 evaluation except for
    genuine ties below ~1e-6 (BASELINE.json's criterion); the 16-bit coarse pass alone is only good to ~1e-5.
    No host synchronisation between the scoring pass and the final indices."""
    assert q.rows_lo is not None and c.rows_lo is not None, "build the stores with exact=True"
    dev = scores.device
    Sq, Sc = scores.shape
    k = min(k, Sc)
    lib = L.lib()
    c_row0, c_len = c.tables()[:2]
    for q0 in range(0, Sq, query_block):
        q1 = min(Sq, q0 + query_block)
        nq = q1 - q0
        sub = q.subset(range(q0, q1)) if (q0, q1) != (0, Sq) else q
        s_row0, s_len, s_item0, n_items = sub.tables()[:4]
        band_idx = torch.empty(nq * Sc, dtype=torch.int32, device=dev)
        counts = torch.empty(nq, dtype=torch.int32, device=dev)
        cands = torch.empty(nq * Sc, 2, dtype=torch.int32, device=dev)
        items = torch.empty(n_items, 6, dtype=torch.int32, device=dev)
        partial = torch.empty(n_items * Sc, dtype=torch.float32, device=dev)
        blk = scores[q0:q1]
        rc = lib.csn_knn_band_select(blk.data_ptr(), scores.stride(0), nq, Sc, k, float(margin), s_row0.data_ptr(),
                                     s_len.data_ptr(), s_item0.data_ptr(), c_row0.data_ptr(), c_len.data_ptr(),
                                     band_idx.data_ptr(), counts.data_ptr(), cands.data_ptr(), items.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_knn_band_select")
        rc = lib.csn_knn_scores_exact(q.rows.data_ptr(), q.rows_lo.data_ptr(), q.rows.shape[0], c.rows.data_ptr(),
                                      c.rows_lo.data_ptr(), c.rows.shape[0], items.data_ptr(), n_items,
                                      cands.data_ptr(), partial.data_ptr(), L.stream_ptr())
        L.check(rc, "csn_knn_scores_exact")
        rc = lib.csn_knn_band_patch(partial.data_ptr(), band_idx.data_ptr(), counts.data_ptr(), s_item0.data_ptr(),
                                    s_len.data_ptr(), nq, Sc, blk.data_ptr(), scores.stride(0), L.stream_ptr())
        L.check(rc, "csn_knn_band_patch")
    return scores


def retrieval_measure(ssa_feats_1: torch.Tensor, ssa_feats_2: torch.Tensor, dtype=torch.float16,
                      eps: float = 1e-12) -> torch.Tensor:
    """get_retrieval_measure (csa_models.py:244-267): (Sq,N,D),(Sc,M,D) fp32 -> (Sq,Sc) fp32."""
    q = build_store(ssa_feats_1, dtype, eps)
    same = ssa_feats_2 is ssa_feats_1
    c = q if same else build_store(ssa_feats_2, dtype, eps)
    return scores_from_stores(q, c)


def knn_graph(ssa_feats_1: torch.Tensor, ssa_feats_2: torch.Tensor, K: int, dtype=torch.float16,
              exact: bool = True) -> torch.Tensor:
    """get_knn_graph (csa_models.py:270-280): indices of the K+1 best candidates per query, sorted
    descending (self included when the two sets coincide).  exact=True re-scores the boundary band so
    that the index sets equal the reference's except for ties below 1e-6."""
    exact = exact and dtype == torch.float16
    q = build_store(ssa_feats_1, dtype, exact=exact)
    c = q if ssa_feats_2 is ssa_feats_1 else build_store(ssa_feats_2, dtype, exact=exact)
    s = scores_from_stores(q, c)
    if exact:
        refine_band(s, q, c, K + 1)
    return topk_rows(s, K + 1)[1]


def cosine_similarity(q: torch.Tensor, k: torch.Tensor, dtype=torch.float16) -> torch.Tensor:
    """HRNetSimCSN.cosine_similarity (hrnet.py:472-490): rows divided by their raw norm (no eps),
    mean over q rows of the max over k rows; q (Lq,D), k (Lk,D) -> 0-d tensor."""
    qs = build_store([q.float()], dtype, 0.0)
    ks = build_store([k.float()], dtype, 0.0)
    return scores_from_stores(qs, ks)[0, 0]
