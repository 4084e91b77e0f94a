"""Seeded synthetic inputs shared by tests, golden-vector generation and bench.py.

Features mimic the reference's `fc_1` activations (post-ReLU, 256-d, N points per shape):
  * iid       : relu(N(0,1))
  * clustered : G latent categories x 32 part prototypes; a point is
                relu(proto[cat][part] + 0.15*shape_jitter[part] + 0.5*N(0,1)); shape s is in
                category s mod G.  Needed for kNN/top-K tests: iid shapes give degenerate,
                near-tied retrieval scores.
Everything is generated on the CPU with torch.Generator so the same seed gives the same bytes in
the build container and on the GPU box (same torch build).
"""
from __future__ import annotations

import torch

D_MODEL = 256
N_POINTS = 10000


def gen(seed: int) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed(int(seed))


def iid_features(g: torch.Generator, *lead: int, n_points: int = N_POINTS, d: int = D_MODEL) -> torch.Tensor:
    """Channel-major MID-FC layout (*lead, d, n_points, 1), fp32."""
    return torch.relu(torch.randn(*lead, d, n_points, 1, generator=g))


def clustered_shapes(seed: int, n_shapes: int, n_points: int = N_POINTS, d: int = D_MODEL,
                     n_categories: int = 16, n_parts: int = 32, first_shape: int = 0) -> torch.Tensor:
    """Row-major point features (n_shapes, n_points, d), fp32; shape ids first_shape.. ."""
    gp = gen(seed)
    protos = torch.randn(n_categories, n_parts, d, generator=gp)
    out = torch.empty(n_shapes, n_points, d)
    for i in range(n_shapes):
        s = first_shape + i
        gs = gen(seed * 1000003 + 17 * s + 1)
        cat = s % n_categories
        jitter = torch.randn(n_parts, d, generator=gs)
        part = torch.randint(0, n_parts, (n_points,), generator=gs)
        noise = torch.randn(n_points, d, generator=gs)
        out[i] = torch.relu(protos[cat][part] + 0.15 * jitter[part] + 0.5 * noise)
    return out


def _uniform(g: torch.Generator, shape, bound: float) -> torch.Tensor:
    return (torch.rand(*shape, generator=g) * 2.0 - 1.0) * bound


def midfc_state(seed: int, n_heads: int, num_classes: int = 15, csa: bool = True, d_model: int = D_MODEL,
                d_k: int = 256, d_v: int = 256) -> dict[str, torch.Tensor]:
    """A full state_dict for csa_models.CrossShapeAt (reference key names) with
    nn.Linear-like magnitudes.  LayerNorm gain/bias are perturbed so that their gradients and the
    gamma/beta code paths are exercised."""
    g = gen(seed)
    hd_k, hd_v = n_heads * d_k, n_heads * d_v
    sd = {
        "fc_1.0.0.weight": _uniform(g, (256, 928, 1, 1), (6.0 / (928 + 256)) ** 0.5),
        "fc_1.0.1.weight": torch.ones(256),
        "fc_1.0.1.bias": torch.zeros(256),
        "fc_1.0.1.running_mean": torch.zeros(256),
        "fc_1.0.1.running_var": torch.ones(256),
        "fc_1.0.1.num_batches_tracked": torch.zeros((), dtype=torch.long),
        "logit.weight": _uniform(g, (num_classes, 256, 1, 1), (6.0 / (256 + num_classes)) ** 0.5),
        "attention.w_qs.weight": _uniform(g, (hd_k, d_model), d_model ** -0.5),
        "attention.w_ks.weight": _uniform(g, (hd_k, d_model), d_model ** -0.5),
        "attention.w_vs.weight": _uniform(g, (hd_v, d_model), d_model ** -0.5),
        "attention.fc.weight": _uniform(g, (d_model, hd_v), hd_v ** -0.5),
        "attention.norm.weight": 1.0 + 0.1 * torch.randn(d_model, generator=g),
        "attention.norm.bias": 0.1 * torch.randn(d_model, generator=g),
    }
    if csa:
        sd.update({
            "compatibility_q.weight": _uniform(g, (256, 256), 256 ** -0.5),
            "compatibility_q.bias": _uniform(g, (256,), 256 ** -0.5),
            "compatibility_k.weight": _uniform(g, (256, 256), 256 ** -0.5),
            "compatibility_k.bias": _uniform(g, (256,), 256 ** -0.5),
        })
    return sd


def mink_state(seed: int, n_head: int = 4, d_model: int = D_MODEL) -> dict[str, torch.Tensor]:
    """Parameters of the MinkowskiNet CSA head (hrnet.py:343,355-356 key names)."""
    g = gen(seed)
    d = d_model
    return {
        "MHA.w_qs.weight": _uniform(g, (d, d), d ** -0.5),
        "MHA.w_ks.weight": _uniform(g, (d, d), d ** -0.5),
        "MHA.w_vs.weight": _uniform(g, (d, d), d ** -0.5),
        "MHA.fc.weight": _uniform(g, (d, d), d ** -0.5),
        "MHA.norm.weight": 1.0 + 0.1 * torch.randn(d, generator=g),
        "MHA.norm.bias": 0.1 * torch.randn(d, generator=g),
        "linear_q.weight": _uniform(g, (d, d), d ** -0.5),
        "linear_k.weight": _uniform(g, (d, d), d ** -0.5),
    }


def csa_batch(seed: int, batch: int, K: int, n_points: int = N_POINTS):
    """(x, x_neighbors) for CrossShapeAt.forward: x (B,256,N,1), x_neighbors (B,K+1,256,N,1) with
    slot 0 = x (features_data_loader.py:133-140)."""
    g = gen(seed)
    nb = iid_features(g, batch, K + 1, n_points=n_points)
    x = nb[:, 0].clone()
    return x, nb


def ragged_lengths(seed: int, n: int, lo: int = 1000, hi: int = 4000) -> list[int]:
    g = gen(seed)
    return [int(v) for v in torch.randint(lo, hi + 1, (n,), generator=g)]
