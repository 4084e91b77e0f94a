"""Recipe that places the UNMODIFIED reference implementation of the hot path under oracle/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot like the built .so files) so that `bench.py --impl reference`
and the cpu_baseline leg time the reference's own PyTorch code on the GPU box's host cores.

The reference is pure Python: "building" it is copying the two modules the path lives in, byte for byte, from where they
lie under /root/reference.  Nothing under oracle/_ref/ is ever committed, imported by the product (csn_b200/), or edited.
Run in the build container (`python oracle/build_ref.py`, also called by __graft_entry__.build()); a no-op elsewhere.
"""
from __future__ import annotations

import hashlib
import shutil
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
FILES = ["MID-FC/csa_models.py", "MinkowskiNet/models/attention.py"]


def build_ref() -> bool:
    """Returns True when oracle/_ref holds the reference files (copied now or earlier)."""
    if not REF.exists():
        return all((OUT / f).exists() for f in FILES)
    manifest = []
    for f in FILES:
        src, dst = REF / f, OUT / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(src.read_bytes()).hexdigest()}  {f}")
    (OUT / "MANIFEST.sha256").write_text("\n".join(manifest) + "\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if build_ref() else "reference not available here")
