"""Reader for the sub-sampled golden vectors written by oracle/make_golden.py."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name: str):
    return np.load(GOLDEN / f"{name}.npz")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative error in the Frobenius norm: |a-b| / |b| (the metric of BASELINE.json's tolerances)."""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def compare_sampled(g, prefix: str, t: torch.Tensor, tol: float, what: str = "") -> float:
    """Check tensor `t` against the golden sample stored under `prefix`; returns the relative error."""
    shape = tuple(int(v) for v in g[f"{prefix}.shape"])
    assert tuple(t.shape) == shape, f"{what or prefix}: shape {tuple(t.shape)} != golden {shape}"
    stride = int(g[f"{prefix}.stride"])
    flat = t.detach().reshape(-1).double().cpu()
    got = flat[::stride]
    want = torch.from_numpy(np.asarray(g[f"{prefix}.values"], dtype=np.float64))
    e = rel_err(got, want)
    assert e <= tol, f"{what or prefix}: sampled rel err {e:.3e} > {tol:.1e}"
    # checksums of the FULL tensor (catch errors outside the sample)
    sumsq = float((flat * flat).sum())
    want_sumsq = float(g[f"{prefix}.sumsq"])
    assert abs(sumsq - want_sumsq) <= max(4 * tol, 1e-6) * max(want_sumsq, 1e-30), \
        f"{what or prefix}: sum of squares {sumsq:.9e} vs golden {want_sumsq:.9e}"
    return e
