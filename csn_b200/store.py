"""GPU-resident feature store + neighbour loader.

The reference's training loop re-loads the per-point features of the query shape and of its K neighbours from
`.npy` files every step (`MID-FC/features_data_loader.py:97-140`, 10 MB per shape, padded to 10 000 points) and moves
the neighbours to the GPU inside the layer (`csa_models.py:216,236`): 328 MB over PCIe per B=8, K=3 step, which is
what bounds the end-to-end step once the layer itself takes ~3.4 ms.  The features are constants of the CSA phase
(the backbone is frozen, `csa_training.py:198-202`), so they can live in HBM: 4 000 shapes x 10.24 MB = 41 GB fp32.

    store = FeatureStore(n_shapes, n_points=10000, device="cuda")
    store.put(ids, feats)                       # once, e.g. while the feature files are read
    x, x_neighbors = store.batch(ids, knn_graph[ids, 1:])   # per step: two device-side gathers, no PCIe traffic
    logits = model(x, mode, x_neighbors)        # the unchanged module surface

`batch` writes into caller-provided (or cached) buffers, so the same static tensors can feed a CUDA graph.
"""
from __future__ import annotations

import torch


class FeatureStore:
    def __init__(self, n_shapes: int, n_points: int = 10000, d_model: int = 256, device="cuda",
                 dtype: torch.dtype = torch.float32):
        self.feats = torch.zeros(n_shapes, d_model, n_points, dtype=dtype, device=device)   # channel-major, as the layer reads it
        self._buf: dict = {}

    @classmethod
    def from_tensor(cls, feats: torch.Tensor) -> "FeatureStore":
        """feats: (S, 256, N) or (S, 256, N, 1), already on the GPU (adopted without a copy when contiguous)."""
        if feats.dim() == 4:
            feats = feats.squeeze(-1)
        s = cls.__new__(cls)
        s.feats = feats.contiguous()
        s._buf = {}
        return s

    @property
    def n_shapes(self) -> int:
        return self.feats.shape[0]

    def put(self, ids, feats: torch.Tensor) -> None:
        """Store features (len(ids), 256, N[, 1]) from host or device memory under the given shape ids."""
        if feats.dim() == 4:
            feats = feats.squeeze(-1)
        idx = torch.as_tensor(ids, dtype=torch.int64, device=self.feats.device)
        self.feats.index_copy_(0, idx, feats.to(self.feats.device, dtype=self.feats.dtype, non_blocking=True))

    def batch(self, ids, nbr_ids, out=None):
        """ids: (B,) shape ids of the queries; nbr_ids: (B, K) ids of their neighbours (host or device ints).
        Returns x (B, 256, N, 1) and x_neighbors (B, K+1, 256, N, 1) in the reference's layout
        (`features_data_loader.py:133-140`); slot 0 of x_neighbors (the query itself) is never read by the layer
        (`csa_models.py:214,234`) and is left unwritten.  `out = (x, x_neighbors)` reuses existing buffers."""
        dev = self.feats.device
        ids = torch.as_tensor(ids, dtype=torch.int64).to(dev, non_blocking=True)
        nbr = torch.as_tensor(nbr_ids, dtype=torch.int64).to(dev, non_blocking=True)
        B, K = nbr.shape
        S, D, N = self.feats.shape
        if out is None:
            key = (B, K)
            if key not in self._buf:
                self._buf[key] = (torch.empty(B, D, N, 1, dtype=torch.float32, device=dev),
                                  torch.empty(B, K + 1, D, N, 1, dtype=torch.float32, device=dev))
            out = self._buf[key]
        x, xn = out
        if self.feats.dtype == torch.float32:
            torch.index_select(self.feats, 0, ids, out=x.view(B, D, N))
            for k in range(K):   # neighbour k+1 of every query: one strided gather per slot
                xn[:, k + 1].view(B, D, N).copy_(self.feats.index_select(0, nbr[:, k]))
        else:
            x.view(B, D, N).copy_(self.feats.index_select(0, ids))
            for k in range(K):
                xn[:, k + 1].view(B, D, N).copy_(self.feats.index_select(0, nbr[:, k]))
        return x, xn
