"""Times alternative tile configurations of the DMMA GEMM on the three projection shapes (dev tool)."""
import ctypes as C, os, torch
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "libgemm_exp.so"))
_p, _i, _ll, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_double
lib.fcest_dgemm.argtypes = [_i, _i, _i, _i, _i, _d, _p, _ll, _ll, _p, _ll, _ll, _d, _p, _ll, _ll, _p, _p, _i, _p, _ll, _p]
lib.fcest_dgemm_workspace_bytes.restype = _ll
lib.fcest_dgemm_workspace_bytes.argtypes = [_i, _i, _i, _ll, _i]
N, R, K = 1200, 2500, 20100
ldk = 20112
dev = "cuda"
Pk = torch.randn(N, ldk, dtype=torch.float64, device=dev); Sk = torch.randn(R, ldk, dtype=torch.float64, device=dev)
g = torch.randn(N, R, dtype=torch.float64, device=dev)
V = torch.empty(N, R, dtype=torch.float64, device=dev); Gk = torch.empty(R, ldk, dtype=torch.float64, device=dev); Tk = torch.empty(N, ldk, dtype=torch.float64, device=dev)
ws = torch.empty(64 << 20, dtype=torch.float64, device=dev)
shapes = {"V=Pk*Sk^T": (1, 1, N, R, K, Pk, ldk, Sk, ldk, V, R), "Gk=g^T*Pk": (0, 0, R, K, N, g, R, Pk, ldk, Gk, ldk),
          "Tk=g*Sk": (1, 0, N, K, R, g, R, Sk, ldk, Tk, ldk)}
def run(sh):
    ak, bk, M_, N_, K_, A, lda, B, ldb, Cc, ldc = sh
    rc = lib.fcest_dgemm(ak, bk, M_, N_, K_, 1.0, A.data_ptr(), lda, 0, B.data_ptr(), ldb, 0, 0.0, Cc.data_ptr(), ldc, 0, None, None, 1,
                         ws.data_ptr(), ws.numel() * 8, None)
    assert rc == 0, rc
for cfg in [0, 3, 4, 5, 6, 7]:
    lib.fcest_dgemm_force_cfg(cfg)
    out = []
    for name, sh in shapes.items():
        for _ in range(2): run(sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(sh)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out.append(f"{name}: {ms:.3f} ms {2.0*N*R*K/ms*1e-9:.2f} TF")
    print("cfg", cfg, " | ".join(out), flush=True)
# reference: cuBLAS on the same shapes
for name, (a, b) in {"V": (Pk[:, :K], Sk[:, :K].T), "Gk": (g.T, Pk[:, :K]), "Tk": (g, Sk[:, :K])}.items():
    for _ in range(2): a @ b
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): a @ b
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
    print("cublas", name, f"{ms:.3f} ms {2.0*N*R*K/ms*1e-9:.2f} TF")
