"""GPU diagnostic for csn_gemm: each case runs in its own process (a trapped kernel poisons the
CUDA context), prints a compact error map on mismatch.  Usage:
    python scripts/gemm_diag.py            # run all cases, one subprocess each
    python scripts/gemm_diag.py CASE_NAME  # run one case in-process
"""
from __future__ import annotations

import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

CASES = {
    # name: (M, N, K, dtype, a_major, b_major, extra)
    "kk_tile": (128, 64, 64, "f16", 0, 0, {}),
    "kk_k256": (128, 128, 256, "f16", 0, 0, {}),
    "kk_256": (256, 256, 256, "f16", 0, 0, {}),
    "kk_ragged": (500, 500, 256, "f16", 0, 0, {}),
    "kk_kragged": (200, 300, 500 - 4, "f16", 0, 0, {}),
    "kk_bf16": (256, 512, 320, "bf16", 0, 0, {}),
    "mnA": (256, 256, 256, "f16", 1, 0, {}),
    "mnB": (256, 256, 256, "f16", 0, 1, {}),
    "mnAB": (384, 192, 448, "f16", 1, 1, {}),
    "mnB_n64": (128, 64, 512, "f16", 0, 1, {}),
    "mnAB_n320": (384, 320, 448, "f16", 1, 1, {}),
    "mnAB_n512": (256, 512, 256, "f16", 1, 1, {}),
    "mnAB_m128": (128, 320, 448, "f16", 1, 1, {}),
    "mnA_ragged": (504, 256, 500 - 4, "f16", 1, 0, {}),
    "out_f16_T": (300, 200, 256, "f16", 0, 0, {"out": "f16", "transposed": True}),
    "out_f32_T": (300, 200, 256, "f16", 0, 0, {"transposed": True}),
    "splitk": (256, 256, 4096, "f16", 1, 1, {"split_k": 8}),
    "batched": (256, 192, 128, "f16", 0, 0, {"batch": (3, 2, 2)}),
    "persist": (2048, 2048, 512, "f16", 0, 0, {}),
    "perf_8k": (8192, 8192, 8192, "f16", 0, 0, {"perf": True}),
    "perf_8k_bf16": (8192, 8192, 8192, "bf16", 0, 0, {"perf": True}),
    "perf_proj": (81920, 768, 256, "f16", 0, 0, {"perf": True}),
    "perf_n256": (37888, 256, 4096, "f16", 0, 0, {"perf": True}),
    "perf_n128": (37888, 128, 4096, "f16", 0, 0, {"perf": True}),
    "perf_n64": (37888, 64, 4096, "f16", 0, 0, {"perf": True}),
    # shapes of the CSA training step (B=8, K=3): rows = 32 slots / 56 blocks x 10240
    "perf_8k_mn": (8192, 8192, 8192, "f16", 1, 1, {"perf": True}),
    # CTA-pair kernel (cta_group::2)
    "pair_small": (512, 512, 256, "f16", 0, 0, {"pair": True}),
    "pair_ragged": (1000, 700, 320, "f16", 0, 0, {"pair": True}),
    "pair_f16out": (768, 512, 512, "bf16", 0, 0, {"pair": True, "out": "bf16"}),
    "pair_8k": (8192, 8192, 8192, "f16", 0, 0, {"pair": True, "perf": True}),
    "pair_qkv": (327680, 768, 256, "f16", 0, 0, {"pair": True, "perf": True, "out": "f16"}),
    "perf_qkv": (327680, 768, 256, "f16", 0, 0, {"perf": True, "out": "f16"}),
    "perf_oproj": (573440, 256, 256, "f16", 0, 0, {"perf": True}),
    "perf_do": (573440, 256, 256, "f16", 0, 1, {"perf": True, "out": "f16"}),
    "perf_wgrad": (768, 256, 327680, "f16", 1, 1, {"perf": True, "split_k": 49}),
    "perf_wgrad_kk": (768, 256, 327680, "f16", 0, 0, {"perf": True, "split_k": 49}),
    "perf_dwo": (256, 256, 573440, "f16", 1, 1, {"perf": True, "split_k": 148}),
    "perf_dwo_kk": (256, 256, 573440, "f16", 0, 0, {"perf": True, "split_k": 148}),
}


def run_case(name: str) -> int:
    import torch
    from csn_b200 import _lib as L

    M, N, K, dt, amn, bmn, ex = CASES[name]
    dtype = torch.float16 if dt == "f16" else torch.bfloat16
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    nb = ex.get("batch", (1, 1, 1))
    nbt = nb[0] * nb[1] * nb[2]
    # logical operands per batch
    A = (torch.randn(nbt, M, K, generator=g) * 0.5).to(dtype)
    B = (torch.randn(nbt, N, K, generator=g) * 0.5).to(dtype)
    big = float(M) * N * K > 2e10
    cm, cn = (256, 256) if big else (M, N)   # big perf cases: check a corner only
    ref = torch.matmul(A[:, :cm].double(), B[:, :cn].double().transpose(1, 2))
    alpha = 0.37
    ref = ref * alpha
    # physical layouts: K-major [nbt*M, K] ; MN-major [nbt*K, M]
    if amn:
        Ap = A.transpose(1, 2).contiguous().view(nbt * K, M).to(dev)
        Am = L.mat(Ap, L.MAJOR_MN, mn_off=(0, 0, 0), k_off=(K, K * nb[0], K * nb[0] * nb[1]))
    else:
        Ap = A.contiguous().view(nbt * M, K).to(dev)
        Am = L.mat(Ap, L.MAJOR_K, mn_off=(M, M * nb[0], M * nb[0] * nb[1]))
    if bmn:
        Bp = B.transpose(1, 2).contiguous().view(nbt * K, N).to(dev)
        Bm = L.mat(Bp, L.MAJOR_MN, k_off=(K, K * nb[0], K * nb[0] * nb[1]))
    else:
        Bp = B.contiguous().view(nbt * N, K).to(dev)
        Bm = L.mat(Bp, L.MAJOR_K, mn_off=(N, N * nb[0], N * nb[0] * nb[1]))
    # NB: batched MN-major with M not the whole inner extent would read neighbours; cases avoid it.
    odt = {"f16": torch.float16, "bf16": torch.bfloat16}.get(ex.get("out", "f32"), torch.float32)
    tr = ex.get("transposed", False)
    split_k = ex.get("split_k", 1)
    if tr:
        D = torch.full((nbt, N, M), float("nan"), dtype=odt, device=dev)
        ld = M
    else:
        D = torch.full((nbt, M, N), float("nan"), dtype=odt, device=dev)
        ld = N
    if split_k > 1:
        D.zero_()
    Dm = L.out(D, ld, transposed=tr, off=(M * N, M * N * nb[0], M * N * nb[0] * nb[1]), accumulate=split_k > 1)
    def run():
        if ex.get("pair"):
            rc = L.lib().csn_gemm_pair(Ap.data_ptr(), Bp.data_ptr(), D.data_ptr(), M, N, K, K, K, N, L.dtype_code(dtype),
                                       L.dtype_code(odt), alpha, L.stream_ptr())
            L.check(rc, "csn_gemm_pair")
        else:
            L.gemm(Am, Bm, Dm, M, N, K, nb=nb, alpha=alpha, split_k=split_k)
    run()
    torch.cuda.synchronize()
    got = D.float().cpu().double()
    if tr:
        got = got.transpose(1, 2)
    got = got[:, :cm, :cn]
    err = (got - ref).abs()
    tol = 2e-2 if odt != torch.float32 or dt == "bf16" else 2e-3
    scale = ref.abs().max().item()
    maxerr = err.max().item()
    nan = torch.isnan(got).sum().item()
    ok = (nan == 0) and maxerr <= tol * max(scale, 1.0)
    print(f"[{name}] M={M} N={N} K={K} {dt} a_mn={amn} b_mn={bmn} {ex} max_err={maxerr:.3e} "
          f"ref_max={scale:.3e} nan={nan} -> {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        e0 = err[0]
        g0 = got[0]
        bad = (e0 > tol * max(scale, 1.0)) | torch.isnan(g0)
        print(f"  bad fraction {bad.double().mean().item():.4f}; first bad idx "
              f"{bad.nonzero()[:5].tolist()}")
        # error map at 8x8 block granularity on the first 128x128 tile
        mm, nn = min(M, 128), min(N, 128)
        blk = bad[:mm, :nn].double()
        rows = []
        for r in range(0, mm, 8):
            rows.append("".join("#" if blk[r:r + 8, c:c + 8].mean() > 0.5 else
                                ("+" if blk[r:r + 8, c:c + 8].any() else ".") for c in range(0, nn, 8)))
        print("  8x8-block error map (first tile):\n    " + "\n    ".join(rows))
        print("  got[0,:4,:8]=\n", g0[:4, :8], "\n  ref[0,:4,:8]=\n", ref[0, :4, :8])
    if ex.get("perf"):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"  perf: {ms:.3f} ms  {2.0 * M * N * K * nbt / ms / 1e9:.1f} TFLOP/s", flush=True)
    return 0 if ok else 1


def main() -> int:
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        return run_case(sys.argv[1])
    names = list(CASES)
    if len(sys.argv) > 1:
        names = [n for n in names if any(a in n for a in sys.argv[1:])]
    fails = 0
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, n], timeout=180, capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                fails += 1
                tail = (r.stderr or "").strip().splitlines()[-6:]
                print(f"[{n}] rc={r.returncode} ({time.time() - t0:.1f}s) stderr tail:\n    " + "\n    ".join(tail))
        except subprocess.TimeoutExpired:
            fails += 1
            print(f"[{n}] TIMEOUT")
        sys.stdout.flush()
    print(f"gemm_diag: {len(names) - fails}/{len(names)} cases OK")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
