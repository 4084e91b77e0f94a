"""ctypes binding of libcsn_b200.so (the C ABI declared in include/csn_b200.h).

There is no CPU fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcsn_b200.so"

CSN_F32, CSN_F16, CSN_BF16 = 0, 1, 2
MAJOR_K, MAJOR_MN = 0, 1

_DTYPE_CODE = {torch.float32: CSN_F32, torch.float16: CSN_F16, torch.bfloat16: CSN_BF16}


class CsnError(RuntimeError):
    pass


class csn_mat(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("major", C.c_int32),
                ("inner", C.c_int64), ("outer", C.c_int64), ("ld", C.c_int64),
                ("mn_off", C.c_int64 * 4), ("k_off", C.c_int64 * 4)]


class csn_out(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("transposed", C.c_int32),
                ("ld", C.c_int64), ("off", C.c_int64 * 4), ("accumulate", C.c_int32),
                ("reserved", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    """The loaded shared library. Raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CsnError(
                f"{LIB_PATH} is missing: build it with `python -m csn_b200.build` "
                "(csn_b200 has no CPU or PyTorch fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        _declare(_lib)
    return _lib


def _declare(L: C.CDLL) -> None:
    L.csn_last_error.restype = C.c_char_p
    L.csn_abi_version.restype = C.c_int
    L.csn_launch_count.restype = C.c_int64
    L.csn_gemm.restype = C.c_int
    L.csn_csa_head_grid.restype = C.c_int
    L.csn_csa_head_grid.argtypes = [C.c_int32, C.c_int32]
    L.csn_gemm.argtypes = [C.POINTER(csn_mat), C.POINTER(csn_mat), C.POINTER(csn_out), C.c_int32,
                           C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_float, C.c_int32,
                           C.c_void_p]
    for name, sig in _EXTRA_SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = sig
    # every kernel-launching entry point goes through a thin wrapper that can time it with CUDA events on
    # the launching stream (bench.py's per-kernel roofline); zero overhead beyond one attribute test otherwise
    for name in list(_EXTRA_SIGNATURES) + ["csn_gemm"]:
        setattr(L, name, _Timed(name, getattr(L, name)))


class _Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *args):
        if _PROFILE is None:
            return self.fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = self.fn(*args)
        e1.record()
        flops = nbytes = 0.0
        if self.name == "csn_gemm":   # (A, B, D, M, N, K, nb[4], alpha, split_k, stream)
            nb = args[6]
            nbt = nb[0] * nb[1] * nb[2] * nb[3]
            M, N, K = args[3], args[4], args[5]
            flops = 2.0 * M * N * K * nbt
            out_b = 4 if args[2]._obj.dtype == CSN_F32 else 2
            nbytes = nbt * (2.0 * (M + N) * K + float(out_b) * M * N)   # operands once + output once (algorithmic)
        elif self.name == "csn_gemm_dual":   # (A0, B0, D0, A1, B1, D1, M, N, K, nb, alpha, stream)
            nb = args[9]
            nbt = nb[0] * nb[1] * nb[2] * nb[3]
            M, N, K = args[6], args[7], args[8]
            flops = 2 * 2.0 * M * N * K * nbt
            nbytes = 2 * nbt * (2.0 * (M + N) * K + 2.0 * M * N)
        elif self.name == "csn_gemm_colbias":   # (A, B, D, M, N, K, alpha, ...)
            M, N, K = args[3], args[4], args[5]
            flops = 2.0 * M * N * K
            out_b = 4 if args[2]._obj.dtype == CSN_F32 else 2
            nbytes = 2.0 * (M + N) * K + float(out_b) * M * N
        elif self.name == "csn_gemm_delta":   # (A, B, dO, lddo, M, N, K, ...): A + O + O_lo in, dO out (16-bit)
            flops = 2.0 * args[4] * args[5] * args[6]
            nbytes = args[4] * (2.0 * args[6] + 3 * 2.0 * args[5])
        elif self.name == "csn_gemm_res_ln":   # (A, B, Z, ldz, M, K, ...), N = 256: A + residual in, Z out
            flops = 2.0 * args[4] * 256 * args[5]
            nbytes = args[4] * (2.0 * args[5] + 4.0 * 256 + 4.0 * 256)
        _PROFILE.append((self.name, e0, e1, flops, nbytes))
        return rc


_PROFILE = None


def profile_begin() -> None:
    """Start recording (name, start event, stop event, flops) for every C-ABI launch."""
    global _PROFILE
    _PROFILE = []


def profile_end() -> dict:
    """Stop recording; returns {entry point: {"ms": total, "launches": n, "flops": algorithmic flops if known}}."""
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out: dict = {}
    for name, e0, e1, fl, nby in rec or []:
        d = out.setdefault(name, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
        d["ms"] += e0.elapsed_time(e1)
        d["launches"] += 1
        d["flops"] += fl
        d["bytes"] += nby
    return out


# name -> argtypes for the remaining entry points (filled in by the sections below)
_EXTRA_SIGNATURES: dict[str, list] = {
    "csn_normalize_rows": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_int32, C.c_void_p],
    "csn_knn_scores": [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                       C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_normalize_rows_split": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p],
    "csn_knn_scores_exact": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                             C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_knn_reduce": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p],
    "csn_pack_rows": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int64,
                      C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                      C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_pack_rows_src16": [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int64,
                            C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                            C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_softmax_fwd": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                        C.c_void_p],
    "csn_softmax_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                        C.c_float, C.c_int32, C.c_void_p],
    "csn_add_ln_fwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                       C.c_int32, C.c_uint32, C.c_float, C.c_void_p],
    "csn_colsum_reduce": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p],
    "csn_gemm_res_ln": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                        C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                        C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p],
    "csn_gemm_colbias": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                         C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p],
    "csn_sgemm_small": [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                        C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p],
    "csn_seg_loss": [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_gemm_delta": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p],
    "csn_gemm_pair": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                      C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_void_p],
    "csn_ln_colsum": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                      C.c_int32, C.c_int32, C.c_void_p],
    "csn_ln_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p],
    "csn_combine_fwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                        C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_block_dot": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_combine_bwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                        C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_attn_fwd": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                     C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                     C.c_void_p, C.c_int32, C.c_uint32, C.c_float, C.c_void_p],
    "csn_attn_delta": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                       C.c_int32, C.c_void_p],
    "csn_attn_bwd_dv": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                        C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                        C.c_void_p, C.c_int32, C.c_uint32, C.c_float, C.c_void_p],
    "csn_attn_bwd_dq": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                        C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                        C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                        C.c_uint32, C.c_float, C.c_void_p],
    "csn_csa_head": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                     C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                     C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_gemm_dual": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                      C.POINTER(C.c_int32), C.c_float, C.c_void_p],
    "csn_grad_unscale": [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
    "csn_attn_bwd_dkv": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                         C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                         C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p],
    "csn_block_add": [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p],
    "csn_ragged_pad": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p],
    "csn_set_drop_epoch": [C.c_uint32, C.c_void_p],
    "csn_segment_mean": [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
    "csn_compat_fanout": [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                          C.c_void_p, C.c_void_p],
    "csn_knn_band_select": [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "csn_knn_band_patch": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                           C.c_int64, C.c_void_p],
    "csn_compat_fwd": [C.c_void_p] * 5 + [C.c_int32, C.c_int32] + [C.c_void_p] * 7,
    "csn_compat_bwd": [C.c_void_p] * 10 + [C.c_int32, C.c_int32] + [C.c_void_p] * 9,
    "csn_topk_rows": [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p],
}


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().csn_last_error().decode()
        raise CsnError(f"{what} failed (rc={rc}): {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().csn_launch_count())


def dtype_code(dt: torch.dtype) -> int:
    return _DTYPE_CODE[dt]


def _i3(vals) -> "C.Array":
    v = [int(x) for x in vals] + [0] * (4 - len(vals))
    return (C.c_int64 * 4)(*v)


def mat(t: torch.Tensor, major: int, mn_off=(), k_off=()) -> csn_mat:
    """Describe a 2-D strided view [outer, inner] (inner contiguous) as a GEMM operand."""
    assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    assert t.dtype in (torch.float16, torch.bfloat16)
    m = csn_mat()
    m.ptr = t.data_ptr()
    m.dtype = dtype_code(t.dtype)
    m.major = major
    m.inner = t.shape[1]
    m.outer = t.shape[0]
    m.ld = t.stride(0)
    m.mn_off = _i3(mn_off)
    m.k_off = _i3(k_off)
    return m


def out(t: torch.Tensor, ld: int, transposed: bool = False, off=(), accumulate: bool = False) -> csn_out:
    assert t.is_cuda
    o = csn_out()
    o.ptr = t.data_ptr()
    o.dtype = dtype_code(t.dtype)
    o.transposed = int(transposed)
    o.ld = ld
    o.off = _i3(off)
    o.accumulate = int(accumulate)
    return o


def gemm(A: csn_mat, B: csn_mat, D: csn_out, M: int, N: int, K: int, nb=(1, 1, 1), alpha: float = 1.0,
         split_k: int = 1) -> None:
    nb3 = (C.c_int32 * 4)(*([int(x) for x in nb] + [1] * (4 - len(nb))))
    rc = lib().csn_gemm(C.byref(A), C.byref(B), C.byref(D), M, N, K, nb3, alpha, split_k, stream_ptr())
    check(rc, "csn_gemm")
