"""C-ABI level parity of the GEMM engine and its fused epilogues against plain fp32/fp64 PyTorch on the same
16-bit operands (the only difference left is the accumulation order): csn_gemm in every operand major, split-K,
csn_gemm_pair (cta_group::2), csn_gemm_res_ln (projection + residual + LayerNorm statistics, csa_models.py:115-118)
and csn_gemm_delta (dO = dZ W_o fused with delta = rowsum(dO o O))."""
import ctypes as C

import pytest
import torch

from csn_b200 import synth

pytestmark = pytest.mark.gpu


def _ops(M, N, K, seed, dt=torch.float16):
    g = synth.gen(seed)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dt).cuda()
    B = (torch.randn(N, K, generator=g) * 0.5).to(dt).cuda()
    return A, B, A.double() @ B.double().t()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm_all_operand_majors(a_mn, b_mn):
    from csn_b200 import _lib as L
    M, N, K = 384, 320, 448
    A, B, ref = _ops(M, N, K, 1)
    At, Bt = A.t().contiguous(), B.t().contiguous()   # kept alive: the descriptors hold raw pointers
    Am = L.mat(At, L.MAJOR_MN) if a_mn else L.mat(A, L.MAJOR_K)
    Bm = L.mat(Bt, L.MAJOR_MN) if b_mn else L.mat(B, L.MAJOR_K)
    D = torch.full((M, N), float("nan"), device="cuda")
    L.gemm(Am, Bm, L.out(D, N), M, N, K, alpha=0.5)
    assert (D.double() - 0.5 * ref).abs().max() < 2e-5 * ref.abs().max()


def test_gemm_split_k_reduce_add_and_bf16():
    from csn_b200 import _lib as L
    M, N, K = 256, 256, 8192
    A, B, ref = _ops(M, N, K, 2, torch.bfloat16)
    D = torch.zeros(M, N, device="cuda")
    L.gemm(L.mat(A, L.MAJOR_K), L.mat(B, L.MAJOR_K), L.out(D, N, accumulate=True), M, N, K, split_k=16)
    assert (D.double() - ref).abs().max() < 2e-5 * ref.abs().max()


@pytest.mark.parametrize("M,N,K,out", [(512, 512, 256, torch.float32), (1000, 700, 320, torch.float32),
                                       (768, 512, 1024, torch.float16)])
def test_gemm_pair_cta_group_2(M, N, K, out):
    from csn_b200 import _lib as L
    A, B, ref = _ops(M, N, K, 3)
    D = torch.full((M, N), float("nan"), dtype=out, device="cuda")
    rc = L.lib().csn_gemm_pair(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, K, K, N, L.dtype_code(A.dtype),
                               L.dtype_code(out), 1.0, L.stream_ptr())
    L.check(rc, "csn_gemm_pair")
    tol = 2e-5 if out == torch.float32 else 1e-3
    assert (D.double() - ref).abs().max() < tol * ref.abs().max()


def test_gemm_res_ln_against_layer_norm():
    """Z = A W^T + residual (read from a channel-major tensor), mean / rstd = LayerNorm statistics (eps 1e-6, biased)."""
    from csn_b200 import _lib as L
    n_blocks, chunk, chunk_pad, n_chunks, n_points = 3, 100, 128, 2, 256   # rows_pad = 256; 200 points used of 256 (row stride must be a 16-byte multiple)
    NP = chunk_pad * n_chunks
    M, K = n_blocks * NP, 256
    g = synth.gen(4)
    A = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    W = (torch.randn(256, K, generator=g) * 0.1).half().cuda()
    res0 = torch.randn(2, 256, n_points, generator=g).cuda()            # two shapes in tensor 0
    res1 = torch.randn(1, 2, 256, n_points, generator=g).cuda()         # (1, K+1, 256, N): slot 1 used
    sel = torch.tensor([0, 1, 0], dtype=torch.int32).cuda()
    row = torch.tensor([256, 256, 0], dtype=torch.int32).cuda()         # block 0 -> res0[1], block 1 -> res1[0,1], block 2 -> res0[0]
    Z = torch.full((M, 256), float("nan"), device="cuda")
    mean = torch.empty(M, device="cuda")
    rstd = torch.empty(M, device="cuda")
    Am, Bm = L.mat(A, L.MAJOR_K), L.mat(W, L.MAJOR_K)
    zbias = torch.randn(n_blocks * n_chunks, 256, generator=g).cuda()   # one row vector per (block, chunk)
    rc = L.lib().csn_gemm_res_ln(C.byref(Am), C.byref(Bm), Z.data_ptr(), 256, M, K, 1.0, res0.data_ptr(), 2 * 256,
                                 res1.data_ptr(), 2 * 256, sel.data_ptr(), row.data_ptr(), n_points, chunk * n_chunks,
                                 NP, chunk_pad, chunk, 1e-6, mean.data_ptr(), rstd.data_ptr(), zbias.data_ptr(), 0, 0.0,
                                 L.stream_ptr())
    L.check(rc, "csn_gemm_res_ln")
    src = [res0[1], res1[0, 1], res0[0]]
    fc = (A.double() @ W.double().t()).view(n_blocks, n_chunks, chunk_pad, 256)
    for b in range(n_blocks):
        for c in range(n_chunks):
            r = src[b][:, c * chunk:(c + 1) * chunk].t().double()           # (chunk, 256)
            z = fc[b, c, :chunk] + r + zbias[b * n_chunks + c].double()
            got = Z.view(n_blocks, n_chunks, chunk_pad, 256)[b, c]
            assert (got[:chunk].double() - z).abs().max() < 1e-5 * z.abs().max()
            assert got[chunk:].abs().max() == 0                               # pad rows are zero
            mu, var = z.mean(dim=1), z.var(dim=1, unbiased=False)
            gm = mean.view(n_blocks, n_chunks, chunk_pad)[b, c, :chunk].double()
            gr = rstd.view(n_blocks, n_chunks, chunk_pad)[b, c, :chunk].double()
            assert (gm - mu).abs().max() < 1e-5
            assert ((gr - (var + 1e-6).rsqrt()).abs() / gr).max() < 1e-5


@pytest.mark.parametrize("n_head,d_head", [(1, 256), (2, 256), (4, 64)])
def test_gemm_delta_against_rowsum(n_head, d_head):
    from csn_b200 import _lib as L
    NP, nblk = 256, 2
    M, HD = nblk * NP, n_head * d_head
    g = synth.gen(5)
    dZ = (torch.randn(M, 256, generator=g) * 0.5).half().cuda()
    Wo = (torch.randn(256, HD, generator=g) * 0.1).half().cuda()          # fc.weight (256, h*d): dO = dZ Wo
    O = torch.randn(M, HD, generator=g).half().cuda()
    O_lo = (torch.randn(M, HD, generator=g) * 0.3).half().cuda()
    dO = torch.full((M, HD), float("nan"), dtype=torch.float16, device="cuda")
    delta = torch.full((nblk * n_head * NP,), float("nan"), device="cuda")
    Am, Bm = L.mat(dZ, L.MAJOR_K), L.mat(Wo, L.MAJOR_MN)
    rc = L.lib().csn_gemm_delta(C.byref(Am), C.byref(Bm), dO.data_ptr(), HD, M, HD, 256, 1.0, O.data_ptr(), O_lo.data_ptr(),
                                HD, delta.data_ptr(), NP, n_head, d_head, L.stream_ptr())
    L.check(rc, "csn_gemm_delta")
    ref = dZ.double() @ Wo.double()
    assert (dO.double() - ref).abs().max() < 2e-3 * ref.abs().max()
    want = (dO.double() * (O.double() + O_lo.double() / 2048.0)).view(nblk, NP, n_head, d_head).sum(-1).permute(0, 2, 1)
    got = delta.view(nblk, n_head, NP).double()
    assert (got - want).abs().max() < 1e-5 * want.abs().max() + 1e-6


def test_gemm_colbias_and_sgemm_small_through_the_c_abi():
    """csn_gemm_colbias (projection with V centred on a per-chunk vector, subtracted in fp32 before rounding) and
    csn_sgemm_small (fp32 contractions with row gathers / transposed operands) against fp64 PyTorch."""
    from csn_b200 import _lib as L
    from csn_b200 import engine as E
    g = synth.gen(14)
    chunk, chunk_pad, n_groups = 100, 128, 5
    M, K, N, col0 = n_groups * chunk_pad, 256, 768, 512
    A = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    W = (torch.randn(N, K, generator=g) * 0.1).half().cuda()
    bias = torch.randn(n_groups, N - col0, generator=g).cuda()
    D = torch.empty(M, N, dtype=torch.float16, device="cuda")
    Am, Bm, Dm = L.mat(A, L.MAJOR_K), L.mat(W, L.MAJOR_K), L.out(D, N)
    rc = L.lib().csn_gemm_colbias(C.byref(Am), C.byref(Bm), C.byref(Dm), M, N, K, 1.0, bias.data_ptr(), N - col0, col0,
                                  chunk_pad, chunk, L.stream_ptr())
    L.check(rc, "csn_gemm_colbias")
    want = (A.double() @ W.double().t()).view(n_groups, chunk_pad, N).cpu().clone()
    want[:, :chunk, col0:] -= bias.double().cpu()[:, None, :]
    got = D.double().view(n_groups, chunk_pad, N).cpu()
    assert (got - want).abs().max() < 2e-3 * want.abs().max()
    # csn_sgemm_small: plain, gathered A rows, and the transposed / gathered "TN" form with accumulation
    a = torch.randn(70, 300, generator=g).cuda()
    b = torch.randn(45, 300, generator=g).cuda()
    d = torch.empty(70, 45, device="cuda")
    E.sgemm_small(a, b, d, 70, 45, 300, alpha=0.5)
    assert torch.allclose(d.double(), 0.5 * a.double() @ b.double().t(), atol=1e-4)
    idx = torch.randint(0, 70, (33,), generator=g).int().cuda()
    d2 = torch.empty(33, 45, device="cuda")
    E.sgemm_small(a, b, d2, 33, 45, 300, a_rows=idx)
    assert torch.allclose(d2.double(), a[idx.long()].double() @ b.double().t(), atol=1e-4)
    # D[m][n] += sum_k A[k][m] * B[rows[k]][n]
    at = torch.randn(50, 70, generator=g).cuda()     # [K][M]
    bt = torch.randn(20, 45, generator=g).cuda()     # table [*][N]
    rows = torch.randint(0, 20, (50,), generator=g).int().cuda()
    d3 = torch.ones(70, 45, device="cuda")
    E.sgemm_small(at, bt, d3, 70, 45, 50, b_rows=rows, trans_a=True, trans_b=True, accumulate=True)
    assert torch.allclose(d3.double(), 1.0 + at.double().t() @ bt[rows.long()].double(), atol=1e-4)
    # few output tiles + long contraction + accumulate: split over grid.z (atomic partial sums)
    at2 = torch.randn(700, 100, generator=g).cuda()
    bt2 = torch.randn(30, 90, generator=g).cuda()
    rows2 = torch.randint(0, 30, (700,), generator=g).int().cuda()
    d4 = torch.full((100, 90), 2.0, device="cuda")
    E.sgemm_small(at2, bt2, d4, 100, 90, 700, b_rows=rows2, trans_a=True, trans_b=True, accumulate=True)
    assert torch.allclose(d4.double(), 2.0 + at2.double().t() @ bt2[rows2.long()].double(), atol=5e-4)


@pytest.mark.parametrize("B,K", [(1, 1), (2, 2), (8, 3)])
def test_compat_glue_kernels_against_torch_fp64(B, K):
    """csn_compat_fwd / csn_compat_bwd (csa_models.py:222-230 incl. the batch-interleaving view) against the
    same lines written with torch in fp64 + autograd."""
    import torch.nn.functional as F
    from csn_b200 import _lib as L
    K1, S = K + 1, B * (K + 1)
    g = synth.gen(100 + B)
    pooled = torch.randn(S, 256, generator=g).cuda()
    Wq, Wk = (torch.randn(256, 256, generator=g) * 0.06).cuda(), (torch.randn(256, 256, generator=g) * 0.06).cuda()
    bq, bk = (torch.randn(256, generator=g) * 0.1).cuda(), (torch.randn(256, generator=g) * 0.1).cuda()
    dcomp = torch.randn(S, generator=g).double().cuda()
    gs = torch.tensor([1.7]).cuda()
    u_q = torch.empty(B, 256, dtype=torch.float64, device="cuda"); u_k = torch.empty(S, 256, dtype=torch.float64, device="cuda")
    nrm = torch.empty(B + S, dtype=torch.float64, device="cuda")
    comp64 = torch.empty(B, K1, dtype=torch.float64, device="cuda"); comp = torch.empty(B, K1, device="cuda")
    rc = L.lib().csn_compat_fwd(pooled.data_ptr(), Wq.data_ptr(), bq.data_ptr(), Wk.data_ptr(), bk.data_ptr(), B, K1, u_q.data_ptr(),
                                u_k.data_ptr(), nrm[:B].data_ptr(), nrm[B:].data_ptr(), comp64.data_ptr(), comp.data_ptr(), L.stream_ptr())
    L.check(rc, "csn_compat_fwd")
    dlin = torch.empty(B + S, 256, dtype=torch.float64, device="cuda")
    gW = torch.empty(2, 256, 256, device="cuda"); gb = torch.empty(2, 256, device="cuda")
    dpool = torch.empty(S, 256, device="cuda"); amax = torch.zeros(1, device="cuda")
    rc = L.lib().csn_compat_bwd(pooled.data_ptr(), Wq.data_ptr(), Wk.data_ptr(), u_q.data_ptr(), u_k.data_ptr(), nrm[:B].data_ptr(),
                                nrm[B:].data_ptr(), comp64.data_ptr(), dcomp.data_ptr(), gs.data_ptr(), B, K1, dlin[:B].data_ptr(),
                                dlin[B:].data_ptr(), gW[0].data_ptr(), gb[0].data_ptr(), gW[1].data_ptr(), gb[1].data_ptr(),
                                dpool.data_ptr(), amax.data_ptr(), L.stream_ptr())
    L.check(rc, "csn_compat_bwd")
    # reference lines, fp64
    p64 = pooled.double().requires_grad_(True)
    w = [t.double().requires_grad_(True) for t in (Wq, bq, Wk, bk)]
    pv = p64.view(B, K1, 256)
    y_q = pv[:, 0]
    y_stack = pv.transpose(0, 1).reshape(K1 * B, 256)                    # [k = 0: b..; k = 1: b..] (:213,220)
    uq = F.normalize(F.linear(y_q, w[0], w[1]), dim=-1)
    uk = F.normalize(F.linear(y_stack, w[2], w[3]), dim=-1).view(B, -1, 256)   # the view of :227
    want = torch.softmax(torch.matmul(uq.unsqueeze(1), uk.permute(0, 2, 1)).squeeze(1), dim=-1)
    want.backward(dcomp.view(B, K1) * 1.7)
    # (the 256-long dot products of the two linears run in fp32; the softmax, normalize and their backward in fp64)
    assert (comp64 - want.detach()).abs().max() < 1e-6 and (comp.double() - want.detach()).abs().max() < 1e-6
    for got, ref in ((gW[0], w[0].grad), (gb[0], w[1].grad), (gW[1], w[2].grad), (gb[1], w[3].grad), (dpool, p64.grad)):
        assert (got.double() - ref).norm() <= 2e-5 * ref.norm() + 1e-12
    assert abs(amax.item() - p64.grad.abs().max().item()) <= 1e-4 * p64.grad.abs().max().item()


def test_gemm_dual_two_problems_in_one_launch():
    """csn_gemm_dual: D0 = A0 B0^T-form (A0 K-major, B0 MN-major) and D1 (A1 = A0^T consumed MN-major, B1 MN-major) written
    into two column ranges of one buffer, batched — the dQ = dS K / dK = dS^T Q pair of the attention backward."""
    from csn_b200 import _lib as L
    g = synth.gen(33)
    nbat, M, N, K = 3, 256, 256, 256          # square dS tiles: M = queries, K = keys for problem 0 and the reverse for problem 1
    dS = (torch.randn(nbat * M, K, generator=g) * 0.5).half().cuda()        # [batch][query][key]
    Kmat = (torch.randn(nbat * K, N, generator=g) * 0.5).half().cuda()      # [batch][key][d]
    Qmat = (torch.randn(nbat * M, N, generator=g) * 0.5).half().cuda()      # [batch][query][d]
    out = torch.full((nbat * M, 3 * N), float("nan"), dtype=torch.float16, device="cuda")
    A0 = L.mat(dS, L.MAJOR_K, mn_off=(M,))
    B0 = L.mat(Kmat, L.MAJOR_MN, k_off=(K,))
    A1 = L.mat(dS, L.MAJOR_MN, k_off=(M,))
    B1 = L.mat(Qmat, L.MAJOR_MN, k_off=(M,))
    D0 = L.out(out[:, :N], 3 * N, off=(M * 3 * N,))
    D1 = L.out(out[:, N:2 * N], 3 * N, off=(M * 3 * N,))
    nb = (C.c_int32 * 4)(nbat, 1, 1, 1)
    rc = L.lib().csn_gemm_dual(C.byref(A0), C.byref(B0), C.byref(D0), C.byref(A1), C.byref(B1), C.byref(D1), M, N, K, nb, 1.0,
                               L.stream_ptr())
    L.check(rc, "csn_gemm_dual")
    dS3, K3, Q3 = dS.double().view(nbat, M, K), Kmat.double().view(nbat, K, N), Qmat.double().view(nbat, M, N)
    want_dq = torch.bmm(dS3, K3).view(nbat * M, N)
    want_dk = torch.bmm(dS3.transpose(1, 2), Q3).view(nbat * K, N)
    assert (out[:, :N].double() - want_dq).abs().max() < 2e-3 * want_dq.abs().max()
    assert (out[:, N:2 * N].double() - want_dk).abs().max() < 2e-3 * want_dk.abs().max()
    assert torch.isnan(out[:, 2 * N:]).all()          # the third column range is untouched


def test_segment_mean_matches_torch():
    """csn_segment_mean against y.mean(dim=0) per shape (hrnet.py:378,388), ragged lengths incl. an empty one."""
    from csn_b200 import _lib as L
    g = torch.Generator().manual_seed(3)
    lens = [5, 0, 1000, 37, 2049]
    x = torch.randn(sum(lens), 256, generator=g).cuda()
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int64).cuda()
    out = torch.empty(len(lens), 256, device="cuda")
    L.check(L.lib().csn_segment_mean(x.data_ptr(), offs.data_ptr(), len(lens), 256, out.data_ptr(), L.stream_ptr()), "seg")
    o = 0
    for i, n in enumerate(lens):
        ref = x[o:o + n].double().mean(dim=0).float() if n else torch.zeros(256, device="cuda")
        assert float((out[i] - ref).abs().max()) < 2e-6
        o += n


def test_block_add_matches_index_add():
    """csn_block_add (residual path of the attention backward) against index_add_ + the power-of-two unscale."""
    from csn_b200 import _lib as L
    g = torch.Generator().manual_seed(5)
    n_src, n_dst, elems = 11, 4, 3 * 256
    src = torch.randn(n_src, elems, generator=g).cuda()
    dst0 = torch.randn(n_dst, elems, generator=g).cuda()
    blk = torch.tensor([0, 1, 2, 3, 0, 0, 2, 1, 0, 3, 3], dtype=torch.int32).cuda()
    amax = torch.tensor([3.0]).cuda()
    for am in (None, amax):
        dst = dst0.clone()
        L.check(L.lib().csn_block_add(src.data_ptr(), blk.data_ptr(), n_src, dst.data_ptr(), n_dst, elems,
                                      am.data_ptr() if am is not None else None, L.stream_ptr()), "block_add")
        want = dst0.double().index_add(0, blk.long(), src.double())
        if am is not None:
            want = want / 2.0 ** float(torch.floor(torch.log2(128.0 / amax.double())))
        assert float((dst.double() - want).abs().max()) < 1e-5


def test_ragged_pad_matches_index_put():
    """csn_ragged_pad: concatenated rows -> zero-padded slots, fp32 and 16-bit copies."""
    from csn_b200 import _lib as L
    g = torch.Generator().manual_seed(9)
    lens, n_pad = [5, 128, 0, 77], 128
    x = torch.randn(sum(lens), 256, generator=g).cuda()
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int64).cuda()
    for dt in (torch.float16, torch.bfloat16):
        o32 = torch.full((len(lens) * n_pad, 256), float("nan"), device="cuda")
        o16 = torch.full((len(lens) * n_pad, 256), float("nan"), device="cuda", dtype=dt)
        amax = torch.zeros(1, device="cuda")
        L.check(L.lib().csn_ragged_pad(x.data_ptr(), offs.data_ptr(), len(lens), n_pad, o32.data_ptr(), o16.data_ptr(),
                                       L.dtype_code(dt), amax.data_ptr(), L.stream_ptr()), "ragged_pad")
        assert float(amax) == float(x.abs().max())
        want = torch.zeros(len(lens), n_pad, 256, device="cuda")
        o = 0
        for s, n in enumerate(lens):
            want[s, :n] = x[o:o + n]
            o += n
        assert torch.equal(o32.view_as(want), want)
        assert torch.equal(o16.view_as(want), want.to(dt))
