"""CPU: host-side tables and the sharding logic of the multi-GPU paths (gloo, world_size 2)."""
import os
import socket

import pytest
import torch

from csn_b200 import engine as E
from csn_b200 import knn, shard


def test_state_dict_keys_match_reference_contract():
    from csn_b200 import midfc, synth
    m = midfc.get_model("csa", 15, 8, 4)
    assert set(m.state_dict().keys()) == set(synth.midfc_state(0, 8, 15).keys())
    assert m.attention.w_qs.weight.shape == (2048, 256)   # d_k = 256 per head
    s = midfc.get_model("ssa", 15, 1)
    assert "compatibility_q.weight" not in s.state_dict()
    with pytest.raises(AttributeError):
        midfc.get_model("nope", 15, 1)


def test_geometry_and_groups():
    g = E.Geometry()
    assert (g.n_points, g.rows_pad, g.kv_len) == (10000, 10240, 500)
    grp = E.Group(n_in=3, n_out=2, blk0=8, q0=0, q_si=0, q_so=4, k0=1, k_si=1, k_so=4, v0=1, v_si=1, v_so=4)
    blocks = list(grp.blocks())
    assert blocks[0] == (8, 0, 1, 1) and blocks[-1] == (13, 4, 7, 7) and len(blocks) == 6


def test_attention_work_tables():
    g = E.Geometry(chunk=500, n_chunks=2, chunk_pad=512)
    grp = E.Group(n_in=2, n_out=1, blk0=0, q0=0, q_si=1, q_so=0, k0=0, k_si=1, k_so=0, v0=0, v_si=1, v_so=0)
    t = E.attn_items([grp], g, 2, 256, "cpu", "fwd")
    assert t.shape == (2 * 2 * 2 * 4, 10)
    assert t[:, 1].tolist()[:4] == [128, 128, 128, 116]          # last tile of a 500-point chunk
    assert int(t[:, 3].min()) == 500 and int(t[:, 5].max()) == 256
    d = E.attn_items([grp], g, 2, 256, "cpu", "dq")
    assert d.shape[1] == 12 and int(d[-1, 7]) == ((1 * 2 + 1) * 2 + 1) * 512 + 384


def test_knn_balance_and_store_subset():
    assert knn._split_for_balance(148, 4000) == 1
    ns = knn._split_for_balance(316, 4000)
    assert (316 * ns) % 148 == 0 or (148 - (316 * ns) % 148) / (316 * ns) < 0.03
    st = knn.ShapeStore(torch.zeros(30, 256, dtype=torch.float16), [0, 10, 20], [10, 10, 10])
    sub = st.subset([2, 0])
    assert sub.row0 == [20, 0] and sub.n_shapes == 2


def test_shard_ranges_cover_everything():
    for n, w in ((4000, 8), (10, 3), (7, 8)):
        spans = [shard.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) kNN: each rank scores its block of queries; rows are gathered in query order
        n_q, n_c, k1 = 7, 5, 3
        gen = torch.Generator().manual_seed(0)
        scores = torch.rand(n_q, n_c, generator=gen)
        lo, hi = shard.shard_range(n_q, rank, world)
        local = scores[lo:hi].topk(k1, dim=-1).indices
        graph = shard.gather_rows(local, n_q, rank, world)
        # (2) training: gradients averaged over ranks with one flat all-reduce
        grads = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))]
        shard.allreduce_mean_(grads, world)
        if rank == 0:
            torch.save({"graph": graph, "ref": scores.topk(k1, dim=-1).indices, "g0": grads[0], "g1": grads[1]}, out)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.equal(r["graph"], r["ref"])
    assert torch.allclose(r["g0"], torch.full((3,), 1.5)) and torch.allclose(r["g1"], torch.full((2, 2), 15.0))


def test_feature_store_gathers_by_shape_id_on_any_device():
    """csn_b200.store.FeatureStore is plain tensor indexing: the gather logic is checked here on the CPU."""
    import torch
    from csn_b200.store import FeatureStore
    S, D, N, B, K = 6, 4, 10, 2, 3
    feats = torch.arange(S * D * N, dtype=torch.float32).view(S, D, N)
    store = FeatureStore(S, n_points=N, d_model=D, device="cpu")
    store.put([3, 0, 5, 1, 4, 2], feats[[3, 0, 5, 1, 4, 2]])
    assert torch.equal(store.feats, feats)
    ids, nbr = [4, 1], [[0, 2, 5], [3, 3, 4]]
    x, xn = store.batch(ids, nbr)
    assert x.shape == (B, D, N, 1) and xn.shape == (B, K + 1, D, N, 1)
    assert torch.equal(x[..., 0], feats[ids])
    for b in range(B):
        for k in range(K):
            assert torch.equal(xn[b, k + 1, ..., 0], feats[nbr[b][k]])
    x2, xn2 = store.batch([0, 0], [[1, 1, 1], [2, 2, 2]])      # buffers are reused for the same (B, K)
    assert x2.data_ptr() == x.data_ptr() and torch.equal(xn2[1, 3, ..., 0], feats[2])


def test_knn_driver_reads_the_reference_feature_layout(tmp_path):
    """csn_b200.knn_driver.FeatureFiles: os.listdir order, (1,256,N,1) files, pad-to-10k by repeating leading points
    (features_data_loader.py:9-48); graph files as csa_training.py:286-290 reads them."""
    import numpy as np
    import torch
    from csn_b200 import knn_driver as D
    root = tmp_path / "train" / "Bed"
    (root / "fc_1").mkdir(parents=True)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((1, 256, 10000, 1)).astype(np.float32)
    b = rng.standard_normal((1, 256, 7000, 1)).astype(np.float32)
    np.save(root / "fc_1" / "a.npy", a)
    np.save(root / "fc_1" / "b.npy", b)
    ff = D.FeatureFiles(str(root))
    assert sorted(ff.files) == ["a.npy", "b.npy"] and len(ff) == 2
    ib = ff.files.index("b.npy")
    fb = ff.load(ib)
    assert fb.shape == (1, 256, 10000, 1)
    assert torch.equal(fb[:, :, :7000], torch.from_numpy(b)) and torch.equal(fb[:, :, 7000:], torch.from_numpy(b[:, :, :3000]))
    batches = list(ff.batches(8))
    assert len(batches) == 1 and batches[0][0].shape == (2, 1, 256, 10000, 1)
    g = np.arange(6, dtype=np.int64).reshape(2, 3)
    D.save_graphs(str(tmp_path / "graphs"), g, g[::-1].copy())
    assert np.array_equal(np.load(tmp_path / "graphs" / "train.npy"), g)
    assert np.array_equal(np.load(tmp_path / "graphs" / "test.npy"), g[::-1])


def _store_worker(rank, world, port, out):
    import torch.distributed as dist
    from csn_b200.store import ShardedFeatureStore
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S, D, N, B, K = 7, 4, 6, 2, 3
        feats = torch.arange(S * D * N, dtype=torch.float32).view(S, D, N)     # the whole collection (for checking)
        store = ShardedFeatureStore(S, n_points=N, d_model=D, device="cpu")
        store.put_local(list(range(store.lo, store.hi)), feats[store.lo:store.hi])
        # rank 0 owns shapes 0..3, rank 1 owns 4..6; neighbours mix local and remote owners, with repeats
        all_ids = [[0, 3], [4, 6]]
        all_nbr = [[[1, 5, 5], [6, 2, 4]], [[0, 5, 3], [3, 3, 6]]]
        x, xn = store.batch(all_ids, all_nbr)
        ok = torch.equal(x[..., 0], feats[all_ids[rank]])
        for b in range(B):
            for k in range(K):
                ok = ok and torch.equal(xn[b, k + 1, ..., 0], feats[all_nbr[rank][b][k]])
        send, recv = store.plan(all_nbr)
        torch.save({"ok": ok, "remote": store.last_remote_blocks, "send": send, "recv": recv}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_feature_store_fetches_neighbours_from_their_owners(tmp_path):
    """csn_b200.store.ShardedFeatureStore over 2 gloo ranks: queries are local, neighbours come from the owning rank,
    every remote block crosses once per step (deduplicated), nothing is sent that the peer already holds."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "store")
    mp.spawn(_store_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert r0["ok"] and r1["ok"]
    assert r0["remote"] == 3 and r0["recv"][1] == [5, 6, 4] and r0["send"][1] == [0, 3]   # rank 0 needs 5, 6, 4 (5 once)
    assert r1["remote"] == 2 and r1["recv"][0] == [0, 3] and r1["send"][0] == [5, 6, 4]


def test_ragged_work_tables_are_longest_first_and_complete(monkeypatch):
    """Work tables of a ragged batch (csn_b200.engine.attn_items): every kind lists the same tiles with and without
    the longest-first order, the order is by streamed length (column 3), and the key-stationary 'dkv' table covers
    every key tile the query-stationary 'dq' table streams."""
    from csn_b200 import engine as E
    lens = (130, 517, 64, 300, 1)
    pairs = ((0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (0, 1), (1, 3), (4, 1))
    n_pad = 640
    geom = E.Geometry(chunk=n_pad, n_chunks=1, chunk_pad=n_pad, kv_chunk=n_pad)
    groups = tuple(E.Group(n_in=1, n_out=1, blk0=j, q0=q, q_si=0, q_so=0, k0=k, k_si=0, k_so=0, v0=k, v_si=0, v_so=0,
                           q_lens=(lens[q],), kv_lens=(lens[k],)) for j, (q, k) in enumerate(pairs))
    h, d = 4, 64
    for kind in ("fwd", "dq", "dkv", "dv"):
        E._ITEM_CACHE.clear()
        monkeypatch.setenv("CSN_ITEM_SORT", "1")
        a = E.attn_items(groups, geom, h, d, "cpu", kind).tolist()
        E._ITEM_CACHE.clear()
        monkeypatch.setenv("CSN_ITEM_SORT", "0")
        b = E.attn_items(groups, geom, h, d, "cpu", kind).tolist()
        assert sorted(map(tuple, a)) == sorted(map(tuple, b)) and len(a) > 0
        assert [r[3] for r in a] == sorted((r[3] for r in a), reverse=True)
    E._ITEM_CACHE.clear()
    monkeypatch.setenv("CSN_ITEM_SORT", "1")
    dq = E.attn_items(groups, geom, h, d, "cpu", "dq").tolist()
    dkv = E.attn_items(groups, geom, h, d, "cpu", "dkv").tolist()
    # per (block, head): query tiles x key tiles of the two tables describe the same rectangle
    n_q_tiles = sum((lens[q] + 127) // 128 for q, _ in pairs) * h
    n_k_tiles = sum((lens[k] + 127) // 128 for _, k in pairs) * h
    assert len(dq) == n_q_tiles and len(dkv) == n_k_tiles
    assert all(r[1] > 0 for r in dq) and all(r[1] > 0 for r in dkv)          # no tile that lies entirely in the padding
    E._ITEM_CACHE.clear()
