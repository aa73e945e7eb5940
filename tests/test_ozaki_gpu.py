"""fcest_dgemm_ozaki (fp64 GEMM emulated on the int8 tensor cores: tcgen05.mma kind::i8, TMEM, TMA) against torch's
float64 matmul.  The error model of the scheme (ozaki.cu header): with ns base-256 digits per operand row,
|C - C_ref|[i][j] <= 4 (ns + 2) K 256^(-ns) (2 max|A_i|) (2 max|B_j|)  (+ fp64 rounding of the reference itself)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _operands(M, N, K, a_kmajor, b_kmajor, seed, scale_rows=False):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, dtype=torch.float64, generator=g)
    B = torch.randn(N, K, dtype=torch.float64, generator=g)
    if scale_rows:   # rows of very different magnitude and a heavy-tailed entry distribution
        A = A * torch.exp(8 * torch.randn(M, 1, dtype=torch.float64, generator=g)) * torch.randn(M, K, dtype=torch.float64, generator=g) ** 3
        B = B * torch.exp(8 * torch.randn(N, 1, dtype=torch.float64, generator=g))
    return A, B


def _run(A, B, a_kmajor, b_kmajor, slices, alpha=1.0, bias=None):
    from fcest_b200 import ops
    M, K = A.shape
    N = B.shape[0]
    Ad = ops.as_pitched((A if a_kmajor else A.t().contiguous()).cuda())
    Bd = ops.as_pitched((B if b_kmajor else B.t().contiguous()).cuda())
    C = ops.pitched(M, N, "cuda")
    C.fill_(float("nan"))
    ops.dgemm_ozaki(a_kmajor, b_kmajor, M, N, K, Ad, Bd, C, alpha=alpha,
                    bias_row=None if bias is None else bias.cuda(), slices=slices)
    torch.cuda.synchronize()
    return C.cpu()


def _check(C, A, B, slices, alpha=1.0, bias=None):
    K = A.shape[1]
    ref = alpha * (A @ B.t())
    if bias is not None:
        ref = ref + bias[:, None]
    bound = 4 * (slices + 2) * K * 256.0 ** (-slices) * (2 * A.abs().amax(1))[:, None] * (2 * B.abs().amax(1))[None, :]
    bound = abs(alpha) * bound + 1e-15 * K ** 0.5 * (A.abs().amax(1)[:, None] * B.abs().amax(1)[None, :]) * abs(alpha) + 1e-300
    bound = bound + 4e-16 * ref.abs()   # rounding of the bias addition / the final scaling
    err = (C - ref).abs()
    assert torch.isfinite(C).all()
    worst = float((err / bound).max())
    assert worst <= 1.0, worst
    return float(err.max() / ref.abs().max())


@pytest.mark.parametrize("M,N,K", [(5, 7, 3), (128, 128, 64), (130, 70, 100), (257, 300, 1000), (64, 520, 129),
                                   (384, 256, 4100)])
@pytest.mark.parametrize("a_kmajor,b_kmajor", [(True, True), (False, False), (True, False), (False, True)])
def test_ozaki_matches_fp64_matmul(cuda, M, N, K, a_kmajor, b_kmajor):
    A, B = _operands(M, N, K, a_kmajor, b_kmajor, seed=M + N + K)
    C = _run(A, B, a_kmajor, b_kmajor, 6)
    rel = _check(C, A, B, 6)
    assert rel < 1e-10
    C = _run(A, B, a_kmajor, b_kmajor, 7)   # 55.9 bits: only the rounding of the two fp64 results is left
    assert _check(C, A, B, 7) < 1e-12


@pytest.mark.parametrize("slices", [2, 3, 4, 5, 6, 7])
def test_ozaki_slice_counts_follow_the_error_model(cuda, slices):
    A, B = _operands(200, 333, 1500, True, True, seed=slices)
    C = _run(A, B, True, True, slices)
    _check(C, A, B, slices)


def test_ozaki_alpha_bias_and_badly_scaled_rows(cuda):
    A, B = _operands(300, 260, 777, True, False, seed=9, scale_rows=True)
    A[17] = 0.0          # an all-zero row of op(A) and of op(B)
    B[5] = 0.0
    bias = torch.randn(300, dtype=torch.float64)
    C = _run(A, B, True, False, 7, alpha=-2.5, bias=bias)
    _check(C, A, B, 7, alpha=-2.5, bias=bias)
    assert torch.equal(C[17], bias[17].expand(260)) and torch.equal(C[:, 5], bias)


def test_ozaki_split_k_path_and_determinism(cuda):
    """Few output tiles and a long contraction: the K range is split over CTAs and summed in fixed order."""
    A, B = _operands(256, 384, 20100, True, True, seed=4)
    C1 = _run(A, B, True, True, 7)
    C2 = _run(A, B, True, True, 7)
    assert torch.equal(C1, C2)
    assert _check(C1, A, B, 7) < 1e-12


def test_ozaki_projection_shapes_match_dmma(cuda):
    """The three projection GEMM shapes of a C2-size model (N=1200, R=225, K=20100) against the native DMMA GEMM."""
    from fcest_b200 import ops
    g = torch.Generator().manual_seed(0)
    N, R, K = 1200, 225, 20100
    Pk = ops.as_pitched(torch.randn(N, K, dtype=torch.float64, generator=g).cuda())
    Sk = ops.as_pitched(torch.randn(R, K, dtype=torch.float64, generator=g).cuda())
    gv = ops.as_pitched(torch.randn(N, R, dtype=torch.float64, generator=g).cuda())
    bias = torch.randn(N, dtype=torch.float64, generator=g).cuda()
    for args, kw in (((True, True, N, R, K, Pk, Sk), {"bias_row": bias}), ((False, False, R, K, N, gv, Pk), {}),
                     ((True, False, N, K, R, gv, Sk), {})):
        M_, N_ = args[2], args[3]
        ref = ops.pitched(M_, N_, "cuda")
        got = ops.pitched(M_, N_, "cuda")
        impl = ops.GEMM_IMPL
        try:
            ops.GEMM_IMPL = "dmma"
            ops.dgemm(*args, ref, **kw)
        finally:
            ops.GEMM_IMPL = impl
        ops.dgemm_ozaki(*args, got, slices=6, **kw)
        rel = float((got - ref).abs().max() / ref.abs().max())
        assert rel < 1e-10, (args[:5], rel)


def test_ozaki_non_finite_operands_give_nan_rows(cuda):
    """A NaN / Inf anywhere in an operand row poisons the matching output row or column (as a native GEMM would), it is
    not silently treated as zero."""
    A, B = _operands(140, 130, 300, True, True, seed=2)
    A[3, 17] = float("nan")
    B[5, 0] = float("inf")
    C = _run(A, B, True, False, 6)
    assert torch.isnan(C[3]).all() and torch.isnan(C[:, 5]).all()
    keep = torch.ones(140, dtype=torch.bool); keep[3] = False
    keepc = torch.ones(130, dtype=torch.bool); keepc[5] = False
    ref = A[keep] @ B[keepc].t()
    assert torch.isfinite(C[keep][:, keepc]).all()
    assert float((C[keep][:, keepc] - ref).abs().max() / ref.abs().max()) < 1e-10


@pytest.mark.timeout(180, method="thread")
@pytest.mark.parametrize("slices,K", [(7, 128), (6, 192), (5, 225), (5, 384), (6, 64), (4, 100)])
def test_ozaki_short_contractions(cuda, slices, K):
    """Short contractions with several work units per CTA: the second accumulator pass is a single stage, issued by ONE of
    the two MMA-issuing warps.  Before both issuers observed every accumulator hand-back this gave wrong results
    (K = 384, 5 digits) or dead-locked (K = 128 / 192 / 225): the other warp skipped a phase of the accumulator-free
    barrier (a parity wait cannot tell two phases behind from done)."""
    from fcest_b200 import ops
    g = torch.Generator().manual_seed(K + slices)
    M, N = 1200, 20100   # 1580 output tiles on 148 SMs
    A = ops.as_pitched(torch.randn(M, K, dtype=torch.float64, generator=g).cuda())
    B = ops.as_pitched(torch.randn(K, N, dtype=torch.float64, generator=g).cuda())
    ref = ops.pitched(M, N, "cuda")
    got = ops.pitched(M, N, "cuda")
    impl = ops.GEMM_IMPL
    try:
        ops.GEMM_IMPL = "dmma"
        ops.dgemm(True, False, M, N, K, A, B, ref)
    finally:
        ops.GEMM_IMPL = impl
    for _ in range(2):
        ops.dgemm_ozaki(True, False, M, N, K, A, B, got, slices=slices)
    torch.cuda.synchronize()
    rel = float((got - ref).abs().max() / ref.abs().max())
    assert rel < {4: 1e-6, 5: 1e-8, 6: 1e-10, 7: 1e-12}[slices], (slices, K, rel)
