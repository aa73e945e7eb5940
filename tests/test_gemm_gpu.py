"""fcest_dgemm (DMMA) against torch.matmul float64 on the same device buffers (tolerance 1e-12 rel)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, dev, g):
    from fcest_b200 import ops
    t = ops.pitched(rows, cols, dev, zero=True)
    t.copy_(torch.randn(rows, cols, dtype=torch.float64, generator=g).to(dev))
    return t


@pytest.mark.parametrize("a_k,b_k", [(True, True), (True, False), (False, True), (False, False)])
# the last two shapes have thousands of output tiles and no split-K (ragged edge tiles, K not a multiple of the k-tile)
@pytest.mark.parametrize("M,N,K", [(7, 9, 5), (64, 64, 16), (130, 257, 100), (200, 1200, 200), (1200, 225, 1035),
                                    (300, 400, 3000), (2500, 4100, 100), (1234, 10001, 333)])
def test_dgemm_layouts(cuda, a_k, b_k, M, N, K):
    from fcest_b200 import ops
    g = torch.Generator().manual_seed(M * 1000 + N + K)
    A = _mk(M, K, cuda, g) if a_k else _mk(K, M, cuda, g)
    B = _mk(N, K, cuda, g) if b_k else _mk(K, N, cuda, g)
    C0 = _mk(M, N, cuda, g)
    bias_r = torch.randn(M, dtype=torch.float64, generator=g).to(cuda)
    bias_c = torch.randn(N, dtype=torch.float64, generator=g).to(cuda)
    opA = A if a_k else A.T
    opB = B.T if b_k else B
    ref = 0.7 * (opA @ opB) + 0.3 * C0 + bias_r[:, None] + bias_c[None, :]
    C = ops.pitched(M, N, cuda)
    C.copy_(C0)
    ops.dgemm(a_k, b_k, M, N, K, A, B, C, alpha=0.7, beta=0.3, bias_row=bias_r, bias_col=bias_c)
    torch.cuda.synchronize()
    err = (C - ref).abs().max().item() / max(ref.abs().max().item(), 1.0)
    assert err < 1e-12, err


def test_dgemm_rejects_misaligned(cuda):
    from fcest_b200 import ops
    A = torch.zeros(8, 9, dtype=torch.float64, device=cuda)  # odd pitch
    B = torch.zeros(8, 9, dtype=torch.float64, device=cuda)
    C = torch.zeros(8, 8, dtype=torch.float64, device=cuda)
    with pytest.raises(RuntimeError):
        ops.dgemm(True, True, 8, 8, 9, A, B, C)


@pytest.mark.parametrize("M", [1, 7, 8, 9, 17, 64, 100, 200, 207, 208, 256])
def test_chol_inv_single(cuda, M):
    """fcest_chol_inv_single (tensor-pipe tile kernel for M <= 207, scalar blocked kernel above) on an
    ill-conditioned Matern-5/2 kernel matrix (random inputs, jitter 1e-6, cond ~ 1e9): backward errors
    |L L^T - K| and |L Li - I| at the level of c M eps, and agreement with torch.linalg to cond-limited 1e-7."""
    import numpy as np
    from fcest_b200 import ops
    rng = np.random.default_rng(M)
    x = np.sort(rng.random(M))[:, None]
    r = np.abs(x - x.T) / 0.3
    K = (1 + np.sqrt(5) * r + 5.0 / 3.0 * r * r) * np.exp(-np.sqrt(5) * r) + 1e-6 * np.eye(M)
    Kt = torch.tensor(K, dtype=torch.float64)
    L_ref = torch.linalg.cholesky(Kt)
    eye = torch.eye(M, dtype=torch.float64)
    Li_ref = torch.linalg.solve_triangular(L_ref, eye, upper=False)
    A = ops.pitched(M, M, cuda)
    A.copy_(Kt.cuda())
    Li = ops.pitched(M, M, cuda)
    info = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.chol_inv_single(A, Li, info)
    assert int(info.item()) == 0
    L, X = A.cpu(), Li.cpu()
    eps = 2.2e-16
    assert (L @ L.T - Kt).abs().max() <= 20.0 * M * eps * float(Kt.abs().max())
    assert (L @ X - eye).abs().max() <= 20.0 * M * eps * float(L.abs().max()) * float(X.abs().max())
    assert (L - L_ref).abs().max() <= 1e-7 * L_ref.abs().max()
    assert (X - Li_ref).abs().max() <= 1e-7 * Li_ref.abs().max()
    assert torch.equal(torch.triu(L, 1), torch.zeros(M, M, dtype=torch.float64))
    assert torch.equal(torch.triu(X, 1), torch.zeros(M, M, dtype=torch.float64))
    # a non-positive pivot is reported LAPACK-style
    bad = Kt.clone()
    bad[M // 2, M // 2] = -1.0
    A.copy_(bad.cuda())
    ops.chol_inv_single(A, Li, info)
    assert int(info.item()) == M // 2 + 1
