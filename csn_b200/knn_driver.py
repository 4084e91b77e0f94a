"""The kNN-graph driver.

The reference's launcher `MID-FC/run_save_knn.py:50-66` runs `python save_knn_graph.py --ssa_logs_dir=... --graphs_dir=...
--partname=... --n_heads=... --num_workers=... --batch_size=... --num_classes=...` for every category, and
`csa_training.py:286-290` loads what it wrote (`train.npy`, `test.npy`: integer arrays (S, K+1), row i = the
retrieval order of shape i, sorted by decreasing score, self included) — but `save_knn_graph.py` itself is not in
the repository.  This module restates it from the graph-refresh step of the trainer
(`csa_training.py:136-166,271-283`) on top of csn_b200's retrieval kernels:

  1. the trained SSA layers are loaded by key (`MID-FC/utils.py:29-39`);
  2. SSA features of every train / test shape (`csa_models.py:282-300`) — kept on the GPU as unit-norm 16-bit rows
     (+ residuals for the exact re-score), 5 MB per shape, instead of a (S, N, 256) fp32 tensor on the host;
  3. small categories: all-pairs scores + top-(K+1) (`csa_models.py:244-280`); big categories
     (`csa_training.py:40`): candidates = shapes nearest to S//10 k-means centres of the amax-pooled features
     (`csa_models.py:302-332`), queries scored against the candidates only, candidate indices mapped back
     (`csa_training.py:138-155`);
  4. `np.save` of the two graphs.

Feature files follow `features_data_loader.py:9-48`: `<root>/fc_1/*.npy` of shape (1, 256, N, 1), listed with
`os.listdir` (the reference's order defines the shape indices), shapes with N < 10 000 padded by repeating their
leading points.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import knn as _knn
from . import midfc

BIG_CLASSES = ("Chair", "Lamp", "StorageFurniture", "Table")     # csa_training.py:40
N_POINTS = 10000


class FeatureFiles:
    """features_data_loader.FeaturesDataset without the labels: file listing and pad-to-10k."""

    def __init__(self, root: str):
        self.dir = os.path.join(root, "fc_1")
        self.files = os.listdir(self.dir)          # the reference indexes shapes in this order

    def __len__(self) -> int:
        return len(self.files)

    def load(self, idx: int) -> torch.Tensor:
        feats = np.load(os.path.join(self.dir, self.files[idx]))          # (1, 256, N, 1)
        if feats.shape[2] < N_POINTS:                                      # features_data_loader.py:37-43
            rem = N_POINTS - feats.shape[2]
            feats = np.concatenate((feats, feats[:, :, :rem, :]), axis=2)
        return torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32))

    def batches(self, batch_size: int):
        """(feats (B, 1, 256, N, 1), None) like the reference's DataLoader batches."""
        for s0 in range(0, len(self), batch_size):
            yield torch.stack([self.load(i) for i in range(s0, min(s0 + batch_size, len(self)))]), None


def ssa_store(model: midfc.CrossShapeAt, files: FeatureFiles, batch_size: int, device, pooled: list = None):
    """SSA features of every shape as a retrieval store on the GPU; optionally also the amax-pooled (S, 256)
    descriptors of csa_models.py:313."""
    dt = midfc._PRECISIONS[model.precision]
    stores = []
    for feats, _ in files.batches(batch_size):
        feats = torch.squeeze(feats.to(device), dim=1)
        with torch.no_grad():
            ssa, _ = model.get_ssa_feats(feats, "test")                    # (B, 256, N, 1)
        rows = ssa.squeeze(-1).permute(0, 2, 1).contiguous()               # (B, N, 256)
        if pooled is not None:
            pooled.append(torch.amax(rows, dim=1))
        stores.append(_knn.build_store(rows, dt, exact=True))
    return _knn.concat_stores(stores)


def _graph(q, c, K: int) -> np.ndarray:
    """top-(K+1) candidate positions per query, boundary band re-scored exactly (index sets equal to the fp32
    reference's except for ties below 1e-6)."""
    s = _knn.scores_from_stores(q, c)
    if q.rows_lo is not None and c.rows_lo is not None:
        _knn.refine_band(s, q, c, K + 1)
    return _knn.topk_rows(s.contiguous(), K + 1)[1].cpu().numpy()


def build_graphs(model, train_root: str, test_root: str, K: int, big: bool, batch_size: int = 8, device="cuda"):
    """(train_graph, test_graph): int64 arrays (S, K+1) as produced by update_knn_graphs (csa_training.py:136-166)."""
    train, test = FeatureFiles(train_root), FeatureFiles(test_root)
    # reject impossible requests BEFORE the SSA features and the scores are computed (minutes of work)
    n_cand = len(train) // 10 if big else len(train)
    if K + 1 > _knn.TOPK_LIMIT:
        raise ValueError(f"K={K}: csn_topk_rows selects at most {_knn.TOPK_LIMIT} entries per row")
    if K + 1 > n_cand:
        raise ValueError(f"K={K}: top-{K + 1} of {n_cand} candidate shapes (torch.topk raises here as well)")
    pooled = [] if big else None
    train_store = ssa_store(model, train, batch_size, device, pooled)
    test_store = ssa_store(model, test, batch_size, device)
    if big:
        from sklearn.cluster import KMeans
        glob = torch.cat(pooled, dim=0).cpu().numpy()
        n_centers = len(glob) // 10
        km = KMeans(n_clusters=n_centers, random_state=0, n_init=10).fit(glob)
        d = ((np.expand_dims(km.cluster_centers_, 1) - glob) ** 2).sum(-1)
        cand = np.argmin(d, axis=-1)
        cand.sort()                                                        # csa_models.py:361
        cand_store = train_store.subset(cand.tolist())
        graphs = [cand[_graph(q, cand_store, K)] for q in (train_store, test_store)]   # candidate positions -> shape indices
        return graphs[0].astype(np.int64), graphs[1].astype(np.int64)
    return _graph(train_store, train_store, K).astype(np.int64), _graph(test_store, train_store, K).astype(np.int64)


def save_graphs(graphs_dir: str, train_graph: np.ndarray, test_graph: np.ndarray) -> None:
    os.makedirs(graphs_dir, exist_ok=True)
    np.save(os.path.join(graphs_dir, "train.npy"), train_graph)           # read back by csa_training.py:286-290
    np.save(os.path.join(graphs_dir, "test.npy"), test_graph)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="save_knn_graph.py of MID-FC (missing from the reference), B200-native")
    ap.add_argument("--ssa_logs_dir", required=True)      # directory with trained_layers.pth
    ap.add_argument("--graphs_dir", required=True)
    ap.add_argument("--partname", required=True)
    ap.add_argument("--n_heads", type=int, default=1)
    ap.add_argument("--num_workers", type=int, default=0)  # accepted for CLI compatibility (files are read in-process)
    ap.add_argument("--batch_size", type=int, default=8)
    ap.add_argument("--num_classes", type=int, required=True)
    ap.add_argument("--K", type=int, default=10)            # run_save_knn.py:34
    ap.add_argument("--testing", action="store_true")
    ap.add_argument("--dataroot", default="data/{}/{}", help="format(split, partname) -> root with fc_1/ (csa_training.py:271-272)")
    args = ap.parse_args(argv)
    dev = torch.device("cuda")
    model = midfc.get_model("ssa", args.num_classes, args.n_heads).to(dev).eval()
    ckpt = torch.load(os.path.join(args.ssa_logs_dir, "trained_layers.pth"), map_location="cpu")
    sd = model.state_dict()
    for k, v in ckpt.items():                                # utils.py:29-39: copy by key
        sd[k].copy_(v)
    model.device = dev
    train_root, test_root = args.dataroot.format("train", args.partname), args.dataroot.format("test", args.partname)
    g_train, g_test = build_graphs(model, train_root, test_root, args.K, args.partname in BIG_CLASSES, args.batch_size, dev)
    save_graphs(args.graphs_dir, g_train, g_test)
    print(f"{args.partname}: train graph {g_train.shape}, test graph {g_test.shape} -> {args.graphs_dir}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
