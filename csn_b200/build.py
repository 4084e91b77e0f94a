"""In-tree build of libcsn_b200.so (hand-written sm_100a kernels + the C ABI of include/csn_b200.h).

nvcc cross-compiles without a GPU.  The shared object is written next to this file so that it
travels with the repository snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libcsn_b200.so"
STAMP = PKG_DIR / ".libcsn_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; csn_b200 has no CPU fallback and cannot be built without it")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                   + [PKG_DIR.parent / "include" / "csn_b200.h"])
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared object. No-op when sources are unchanged."""
    if not force and is_current():
        return LIB_PATH
    objs = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    for src in sources():
        obj = build_dir / (src.stem + ".o")
        objs.append(obj)
        cmd = [_nvcc(), *compile_flags, "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src.name}\n{out}")
        failed |= p.returncode != 0
    (build_dir / "nvcc.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see csn_b200/build/nvcc.log")
    link = [_nvcc(), "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
            *map(str, objs), "-o", str(LIB_PATH)]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libcsn_b200.so failed")
    STAMP.write_text(_digest())
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
