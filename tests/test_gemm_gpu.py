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
@pytest.mark.parametrize("M,N,K", [(7, 9, 5), (64, 64, 16), (130, 257, 100), (200, 1200, 200), (1200, 225, 1035),
                                    (300, 400, 3000)])
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
