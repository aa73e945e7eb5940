"""Parity of the CUDA Wishart path (through the C ABI) against the CPU oracle on seeded inputs.

Tolerances are the ones BASELINE.json's north_star states: ELBO 1e-8 relative; gradients are held to
1e-7 relative-to-max (they feed the same ELBO); predicted covariances / correlations 1e-6.
"""
import numpy as np
import pytest
import torch

from oracle import wishart_oracle as wo

pytestmark = pytest.mark.gpu


def _build(kind, N, D, S, M, a_option="train_full_matrix", seed=0):
    from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess
    x, y, eps, p = wo.make_inputs(N, D, S, M=M, seed=seed, a_option=a_option)
    if kind == "svgp":
        m = SparseVariationalWishartProcess(D, p["Z"].numpy(), num_mc_samples=S, A_scale_matrix_option=a_option,
                                            verbose=False)
        m.inducing_variable.Z.assign(p["Z"])
    else:
        m = VariationalWishartProcess(x.numpy(), y.numpy(), num_mc_samples=S, A_scale_matrix_option=a_option)
    m.q_mu.assign(p["q_mu"])
    m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"])
    m.likelihood.additive_part.assign(p["lam"])
    m.likelihood.additive_part.trainable = True
    return m, x, y, eps, p


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


CASES = [("svgp", 40, 3, 4, 12), ("svgp", 33, 5, 3, 17), ("svgp", 96, 15, 3, 40), ("vgp", 30, 3, 5, None),
         ("vgp", 41, 4, 2, None), ("svgp", 64, 18, 2, 24)]


@pytest.mark.parametrize("kind,N,D,S,M", CASES)
def test_elbo_and_grad_parity(cuda, kind, N, D, S, M):
    m, x, y, eps, p = _build(kind, N, D, S, M)
    elbo_ref, g_ref = wo.elbo_and_grad(kind, p, x, y, eps)
    data = (x.numpy(), y.numpy()) if kind == "svgp" else None
    elbo = m.elbo_and_grad(data, epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < 1e-8, (elbo, elbo_ref)
    # forward-only path gives the same number
    elbo_f = m.elbo(data, epsilon=eps.cuda()).item()
    assert abs(elbo_f - elbo_ref) / abs(elbo_ref) < 1e-8
    got = {"q_mu": m.q_mu.grad, "q_sqrt": m.q_sqrt.grad, "A": m.likelihood.A_scale_matrix.grad,
           "lam": m.likelihood.additive_part.grad, "u_var": m.kernel.variance.grad,
           "u_ls": m.kernel.lengthscales.grad}
    if kind == "svgp":
        got["Z"] = m.inducing_variable.Z.grad
    for k, v in got.items():
        r = _rel(v.cpu().numpy().reshape(-1), g_ref[k].numpy().reshape(-1))
        assert r < 1e-7, (k, r)


@pytest.mark.parametrize("a_option", ["identity", "scaled", "train_diagonal"])
def test_vector_A_options(cuda, a_option):
    m, x, y, eps, p = _build("svgp", 40, 4, 3, 10, a_option=a_option)
    elbo_ref, g_ref = wo.elbo_and_grad("svgp", p, x, y, eps)
    elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < 1e-8
    assert _rel(m.likelihood.A_scale_matrix.grad.cpu().numpy(), g_ref["A"].numpy()) < 1e-7
    assert _rel(m.q_sqrt.grad.cpu().numpy(), g_ref["q_sqrt"].numpy()) < 1e-7


def test_reference_init_point(cuda):
    """At the reference initialisation (A = I, q_mu = 0, q_sqrt = 1e-3 I) Sigma is diagonal."""
    from fcest_b200.models import SparseVariationalWishartProcess
    x, y, eps, p = wo.make_inputs(50, 6, 5, M=20, perturbed=False)
    m = SparseVariationalWishartProcess(6, p["Z"].numpy(), num_mc_samples=5, verbose=False)
    elbo_ref, g_ref = wo.elbo_and_grad("svgp", p, x, y, eps)
    elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < 1e-8
    assert _rel(m.q_sqrt.grad.cpu().numpy(), g_ref["q_sqrt"].numpy()) < 1e-7


def test_predict_parity(cuda):
    for kind, N, D, S, M in [("svgp", 40, 3, 4, 12), ("vgp", 30, 3, 5, None), ("svgp", 50, 10, 2, 16)]:
        m, x, y, eps, p = _build(kind, N, D, S, M)
        xn = torch.linspace(-0.1, 1.1, 23, dtype=torch.float64)[:, None]
        if kind == "svgp":
            fm_ref, fv_ref = wo.svgp_predict_f(xn, p)
        else:
            fm_ref, fv_ref = wo.vgp_predict_f(xn, x, p)
        fm, fv = m.predict_f(xn.numpy())
        assert _rel(fm.cpu().numpy(), fm_ref.numpy()) < 1e-9
        assert _rel(fv.cpu().numpy(), fv_ref.numpy()) < 1e-9
        g = torch.Generator().manual_seed(7)
        S_new = 37
        eps_new = torch.randn(S_new, 23, D, D, dtype=torch.float64, generator=g)
        cm_ref, cs_ref = wo.predict_cov(fm_ref, fv_ref, p["A"], p["lam"], eps_new, D)
        rm_ref, rs_ref = wo.predict_corr(fm_ref, fv_ref, p["A"], p["lam"], eps_new, D)
        cm, cs = m.predict_cov(xn.numpy(), S_new, epsilon=eps_new.cuda())
        rm, rs = m.predict_corr(xn.numpy(), S_new, epsilon=eps_new.cuda())
        for a, b in [(cm, cm_ref), (cs, cs_ref), (rm, rm_ref), (rs, rs_ref)]:
            assert np.abs(a.cpu().numpy() - b.numpy()).max() < 1e-6 * max(1.0, float(b.abs().max()))
        smp = m._get_cov_samples(xn.numpy(), S_new, epsilon=eps_new.cuda())
        ref = wo.cov_samples(fm_ref, fv_ref, p["A"], p["lam"], eps_new, D)
        assert _rel(smp.cpu().numpy(), ref.numpy()) < 1e-10


def test_log_prob_identities(cuda):
    """The reference's own likelihood test identities (tests/fcest/models/test_likelihoods.py:54-76),
    exercised through _log_prob: value equals the dense MVN log density computed with numpy."""
    from fcest_b200.models.likelihoods import WishartProcessLikelihood
    rng = np.random.default_rng(3)
    S, N, D = 2, 6, 4
    lik = WishartProcessLikelihood(D, num_mc_samples=S, verbose=False)
    F = rng.standard_normal((S, N, D, D))
    y = rng.standard_normal((N, D))
    got = lik._log_prob(None, torch.tensor(F).cuda(), y).cpu().numpy()
    A = np.eye(D)
    lam = np.log1p(np.exp(lik.additive_part.numpy()))
    ref = np.zeros(N)
    for s in range(S):
        for n in range(N):
            G = A * F[s, n]
            Sig = G @ G.T + np.diag(lam)
            sign, logdet = np.linalg.slogdet(Sig)
            ref[n] += (-0.5 * D * np.log(2 * np.pi) - 0.5 * logdet - 0.5 * y[n] @ np.linalg.solve(Sig, y[n])) / S
    assert np.abs(got - ref).max() < 1e-10


def test_cholesky_failure_raises(cuda):
    from fcest_b200.models.likelihoods import WishartProcessLikelihood
    lik = WishartProcessLikelihood(3, num_mc_samples=1, verbose=False)
    lik.additive_part.assign(np.full(3, -800.0))  # softplus -> 0, Sigma = G G^T singular for F = 0
    with pytest.raises(RuntimeError):
        lik._log_prob(None, torch.zeros(1, 2, 3, 3, dtype=torch.float64).cuda(), np.zeros((2, 3)))


def test_adam_matches_keras_formula(cuda):
    from fcest_b200 import ops
    g = torch.Generator().manual_seed(0)
    theta = torch.randn(1000, dtype=torch.float64, generator=g)
    m = torch.zeros_like(theta)
    v = torch.zeros_like(theta)
    th_d, m_d, v_d = theta.cuda(), m.cuda(), v.cuda()
    for t in range(1, 4):
        grad_elbo = torch.randn(1000, dtype=torch.float64, generator=g)
        theta, m, v = wo.keras_adam_step(theta, -grad_elbo, m, v, t)
        ops.adam_step(th_d, grad_elbo.cuda(), m_d, v_d, t)
    assert (th_d.cpu() - theta).abs().max() < 1e-15


def test_run_adam_reference_smoke(cuda):
    """Reference tests/fcest/models/test_wishart_process.py:23-80: 3 iterations return 3 ELBOs."""
    from fcest_b200.helpers.inference import run_adam
    from fcest_b200.models import SparseVariationalWishartProcess, VariationalWishartProcess
    rng = np.random.default_rng(0)
    x = np.arange(7, dtype=np.float64)[:, None]
    y = rng.random((7, 2))
    m = SparseVariationalWishartProcess(2, x.copy(), nu=2, verbose=False)
    logf = run_adam("SVWP", m, iterations=3, log_interval=1, data=(x, y))
    assert type(logf) is list and len(logf) == 3
    m = VariationalWishartProcess(x, y, nu=2)
    logf = run_adam("VWP", m, iterations=3, log_interval=1)
    assert type(logf) is list and len(logf) == 3 and all(np.isfinite(logf))


def test_adam_trajectory_parity(cuda):
    """Three Adam steps with injected noise follow the oracle's Keras-Adam trajectory."""
    from fcest_b200.helpers.inference import KerasAdam
    m, x, y, eps, p = _build("svgp", 30, 3, 3, 9)
    names = list(wo.TRAINABLE)
    state = {k: (torch.zeros_like(p[k]), torch.zeros_like(p[k])) for k in names}
    opt = KerasAdam(m.trainable_parameters)
    for t in range(1, 4):
        _, g = wo.elbo_and_grad("svgp", p, x, y, eps)
        for k in names:
            mm, vv = state[k]
            p[k], mm, vv = wo.keras_adam_step(p[k], -g[k].reshape(p[k].shape), mm, vv, t)
            state[k] = (mm, vv)
        m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda())
        opt.step()
    assert _rel(m.q_mu.unconstrained.cpu().numpy(), p["q_mu"].numpy()) < 1e-9
    assert _rel(m.q_sqrt.unconstrained.cpu().numpy(), p["q_sqrt"].numpy()) < 1e-9
    assert _rel(m.likelihood.A_scale_matrix.unconstrained.cpu().numpy(), p["A"].numpy()) < 1e-9
    assert abs(m.kernel.lengthscales.unconstrained.item() - p["u_ls"].item()) < 1e-9


def test_large_D_path(cuda):
    """D > 80 uses the global-scratch variants of the likelihood / predict kernels."""
    kind, N, D, S, M = "svgp", 12, 96, 2, 8
    m, x, y, eps, p = _build(kind, N, D, S, M)
    elbo_ref, g_ref = wo.elbo_and_grad(kind, p, x, y, eps)
    elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    assert abs(elbo - elbo_ref) / abs(elbo_ref) < 1e-8, (elbo, elbo_ref)
    for k, v in {"q_sqrt": m.q_sqrt.grad, "A": m.likelihood.A_scale_matrix.grad,
                 "lam": m.likelihood.additive_part.grad, "q_mu": m.q_mu.grad}.items():
        assert _rel(v.cpu().numpy().reshape(-1), g_ref[k].numpy().reshape(-1)) < 1e-7, k
    xn = torch.linspace(0, 1, 5, dtype=torch.float64)[:, None]
    fm_ref, fv_ref = wo.svgp_predict_f(xn, p)
    g = torch.Generator().manual_seed(1)
    eps_new = torch.randn(6, 5, D, D, dtype=torch.float64, generator=g)
    cm_ref, cs_ref = wo.predict_cov(fm_ref, fv_ref, p["A"], p["lam"], eps_new, D)
    rm_ref, rs_ref = wo.predict_corr(fm_ref, fv_ref, p["A"], p["lam"], eps_new, D)
    cm, cs = m.predict_cov(xn.numpy(), 6, epsilon=eps_new.cuda())
    rm, rs = m.predict_corr(xn.numpy(), 6, epsilon=eps_new.cuda())
    for a, b in [(cm, cm_ref), (cs, cs_ref), (rm, rm_ref), (rs, rs_ref)]:
        assert np.abs(a.cpu().numpy() - b.numpy()).max() < 1e-6 * max(1.0, float(b.abs().max()))


def test_pd_check_large_and_truth_table(cuda):
    """are_all_positive_definite: the reference's truth table (tests/fcest/helpers/test_array_operations.py:16-65)
    and a D = 200 batch (global-scratch path)."""
    from fcest_b200.helpers.array_operations import are_all_positive_definite
    assert not are_all_positive_definite(np.array([[[1.0, 2.0], [2.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]]]))
    assert not are_all_positive_definite(np.array([[[1.0, 0.5], [0.4, 1.0]]]))
    assert are_all_positive_definite(np.array([[[2.0, 0.5], [0.5, 1.0]], [[1.0, 0.0], [0.0, 1.0]]]))
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 200, 220))
    spd = a @ np.swapaxes(a, 1, 2)
    assert are_all_positive_definite(spd)
    spd[1, 5, 5] = -1.0
    assert not are_all_positive_definite(spd)


def test_time_sharded_matches_single_gpu(cuda):
    """TimeShardedSVWP at world size 3, emulated on one GPU: pass 1 records what every rank would
    contribute to each all-reduce, pass 2 replays the ranks with the recorded sums.  ELBO and every
    gradient must equal the unsharded evaluation (sums over time points re-associated: 1e-12)."""
    from fcest_b200.parallel import TimeShardedSVWP
    kind, N, D, S, M = "svgp", 50, 6, 3, 14
    m, x, y, eps, p = _build(kind, N, D, S, M)
    ref_elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    ps = [m.q_mu, m.q_sqrt, m.kernel.variance, m.kernel.lengthscales, m.likelihood.A_scale_matrix,
          m.likelihood.additive_part, m.inducing_variable.Z]
    ref = [q.grad.clone() for q in ps]
    world = 3
    recorded = [[] for _ in range(world)]
    for r in range(world):
        sh = TimeShardedSVWP(m, world=world, rank=r, reduce_=lambda t, r=r: recorded[r].append(t.clone()))
        sh.elbo_and_grad((x.cuda(), y.cuda()), epsilon=eps.cuda())
    sums = [sum(recorded[r][c] for r in range(world)) for c in range(len(recorded[0]))]
    for r in range(world):
        calls = iter(sums)
        sh = TimeShardedSVWP(m, world=world, rank=r, reduce_=lambda t: t.copy_(next(calls)))
        elbo = sh.elbo_and_grad((x.cuda(), y.cuda()), epsilon=eps.cuda()).item()
        assert abs(elbo - ref_elbo) <= 1e-12 * abs(ref_elbo)
        for q, g in zip(ps, ref):
            assert _rel(q.grad.cpu().numpy().reshape(-1), g.cpu().numpy().reshape(-1)) < 1e-11, q.name


@pytest.mark.parametrize("world", [2, 3])
def test_latent_sharded_matches_single_gpu(cuda, world):
    """LatentShardedSVWP (conditional sharded by latent, likelihood by time), emulated on one GPU: the ranks'
    phases run in lock-step and the two all-to-alls / the all-reduce are list transpositions / sums.  ELBO and
    every gradient must equal the unsharded evaluation (sums re-associated: 1e-11)."""
    from fcest_b200.parallel import LatentShardedSVWP
    kind, N, D, S, M = "svgp", 50, 5, 3, 14       # R = 25: uneven, odd latent blocks
    m, x, y, eps, p = _build(kind, N, D, S, M)
    ref_elbo = m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda()).item()
    small = [m.kernel.variance, m.kernel.lengthscales, m.likelihood.A_scale_matrix, m.likelihood.additive_part,
             m.inducing_variable.Z]
    ref_small = [q.grad.clone() for q in small]
    ref_qmu, ref_qsqrt = m.q_mu.grad.clone(), m.q_sqrt.grad.clone()
    xs, ys, ec = x.cuda(), y.cuda(), eps.cuda()
    ranks = [LatentShardedSVWP(m, world=world, rank=r) for r in range(world)]
    sends = [sh.forward_local(xs) for sh in ranks]                                   # sends[p][q]
    recvs = [[sends[p][q] for p in range(world)] for q in range(world)]             # all-to-all
    for q, sh in enumerate(ranks):
        assert [tuple(t.shape) for t in sh.recv_buffers_forward(xs.device)] == [tuple(t.shape) for t in recvs[q]]
    sends = [sh.likelihood_local(xs, ys, recvs[r], epsilon=ec) for r, sh in enumerate(ranks)]
    recvs = [[sends[p][q] for p in range(world)] for q in range(world)]
    for q, sh in enumerate(ranks):
        assert [tuple(t.shape) for t in sh.recv_buffers_backward(xs.device)] == [tuple(t.shape) for t in recvs[q]]
    flats = [sh.backward_local(recvs[r]) for r, sh in enumerate(ranks)]
    total = sum(flats)                                                               # all-reduce
    for r, sh in enumerate(ranks):
        elbo = sh.finish(total.clone()).item()
        assert abs(elbo - ref_elbo) <= 1e-12 * abs(ref_elbo)
        for q, g in zip(small, ref_small):
            assert _rel(q.grad.cpu().numpy().reshape(-1), g.cpu().numpy().reshape(-1)) < 1e-11, q.name
        lat = sh.latents
        assert _rel(sh.q_mu_grad_local.cpu().numpy(), ref_qmu[:, lat].cpu().numpy()) < 1e-11
        assert _rel(sh.q_sqrt_grad_local.cpu().numpy(), ref_qsqrt[lat].cpu().numpy()) < 1e-11
    assert sum(len(b) for b in ranks[0].latent_blocks) == D * D


def test_latent_sharded_adam_matches_single_gpu_trajectory(cuda):
    """Three Adam steps with sharded q_sqrt / q_mu optimiser state (two emulated ranks in lock-step) against the
    unsharded KerasAdam trajectory on the same noise."""
    from fcest_b200.helpers.inference import KerasAdam
    from fcest_b200.parallel import LatentShardedSVWP
    kind, N, D, S, M = "svgp", 40, 4, 3, 10
    world, steps = 2, 3
    m_ref, x, y, eps, p = _build(kind, N, D, S, M)
    m_sh, *_ = _build(kind, N, D, S, M)
    xs, ys = x.cuda(), y.cuda()
    g = torch.Generator().manual_seed(9)
    noise = [torch.randn(S, N, D * D, dtype=torch.float64, generator=g).cuda() for _ in range(steps)]
    opt = KerasAdam(m_ref.trainable_parameters)
    ranks = [LatentShardedSVWP(m_sh, world=world, rank=r) for r in range(world)]
    for it in range(steps):
        m_ref.elbo_and_grad((xs, ys), epsilon=noise[it])
        opt.step()
        sends = [sh.forward_local(xs) for sh in ranks]
        recvs = [[sends[p_][q] for p_ in range(world)] for q in range(world)]
        sends = [sh.likelihood_local(xs, ys, recvs[r], epsilon=noise[it]) for r, sh in enumerate(ranks)]
        recvs = [[sends[p_][q] for p_ in range(world)] for q in range(world)]
        flats = [sh.backward_local(recvs[r]) for r, sh in enumerate(ranks)]
        total = sum(flats)
        for r, sh in enumerate(ranks):
            sh.finish(total.clone())
        for r, sh in enumerate(ranks):
            sh.adam_step(update_small=(r == 0))     # the emulated ranks share one model object
    for a, b in zip(m_ref.parameters, m_sh.parameters):
        assert _rel(b.unconstrained.cpu().numpy().reshape(-1), a.unconstrained.cpu().numpy().reshape(-1)) < 1e-9, a.name


@pytest.mark.parametrize("D", [7, 8, 33, 50, 56, 57])
def test_predict_moments_all_tile_counts(cuda, D):
    """fcest_cov_moments: register-tiled kernel (D <= 56, every row-block count incl. ragged warps) and the
    shared-memory kernel just above it, in two chunks (running Welford state) and with raw samples."""
    from fcest_b200 import ops
    rng = np.random.default_rng(D)
    N, S1, S2 = 5, 4, 3
    R = D * D
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    fm, fv = t(0.3 * rng.standard_normal((N, R))), t(0.05 + rng.random((N, R)))
    A, lam = t(np.eye(D) + 0.2 * rng.standard_normal((D, D))), t(wo.inv_softplus(0.01) + 0.3 * rng.standard_normal(D))
    eps = t(rng.standard_normal((S1 + S2, N, D, D)))
    for corr, ref_fn in ((0, wo.predict_cov), (1, wo.predict_corr)):
        m_ref, s_ref = ref_fn(fm, fv, A, lam, eps, D)
        fm_d, fv_d = ops.as_pitched(fm.cuda()), ops.as_pitched(fv.cuda())
        mean = torch.empty((N, D, D), dtype=torch.float64, device=cuda)
        m2 = torch.empty_like(mean); om = torch.empty_like(mean); os_ = torch.empty_like(mean)
        smp = torch.empty((S1 + S2, N, D, D), dtype=torch.float64, device=cuda)
        nscr = int(ops._lib.load_library().fcest_cov_moments_scratch_bytes(D, D))
        scr = torch.empty(max(nscr // 8, 1), dtype=torch.float64, device=cuda)
        e_d = eps.reshape(S1 + S2, N, R).cuda().contiguous()
        A_d, lam_d = A.cuda(), lam.cuda()  # keep the device copies alive across the asynchronous launches
        for (lo, hi, last) in ((0, S1, 0), (S1, S1 + S2, 1)):
            ops.call("fcest_cov_moments", ops.ptr(fm_d), ops.ptr(fv_d), ops.ld(fm_d), ops.ptr(e_d[lo:hi].contiguous()),
                     ops.ptr(A_d), 0, ops.ptr(lam_d), hi - lo, N, D, D, corr, lo, ops.ptr(mean), ops.ptr(m2),
                     ops.ptr(smp[lo:hi]), last, ops.ptr(om), ops.ptr(os_), ops.ptr(scr) if nscr else None)
        assert np.abs(om.cpu().numpy() - m_ref.numpy()).max() < 1e-9 * max(1.0, float(m_ref.abs().max()))
        assert np.abs(os_.cpu().numpy() - s_ref.numpy()).max() < 1e-9 * max(1.0, float(s_ref.abs().max()))
        ref_s = wo.cov_samples(fm, fv, A, lam, eps, D)
        if corr:
            ref_s = wo.convert_tensor_to_correlation(ref_s)
        assert _rel(smp.cpu().numpy(), ref_s.numpy()) < 1e-10


@pytest.mark.parametrize("kind", ["svgp", "vgp"])
def test_cuda_graph_training_step_matches_eager(cuda, kind):
    """GraphedTrainingStep (eager first call, captured second, replayed afterwards) follows exactly the
    trajectory of the eager loop `elbo_and_grad + KerasAdam.step` fed with the same Philox draws."""
    from fcest_b200 import ops
    from fcest_b200.helpers.inference import GraphedTrainingStep, KerasAdam
    N, D, S, M = 40, 5, 3, 12
    steps = 5
    traj = []
    for graphed in (False, True):
        m, x, y, eps, p = _build(kind, N, D, S, M if kind == "svgp" else None)
        data = (x.numpy(), y.numpy()) if kind == "svgp" else None
        m.likelihood.seed(77)
        opt = KerasAdam(m.trainable_parameters)
        elbos = []
        if graphed:
            g = GraphedTrainingStep(m, data, opt)
            for _ in range(steps):
                elbos.append(float(g().item()))
            g.check()
        else:
            xd, yd = m._data(data)
            e = torch.empty((S, yd.shape[0], D * D), dtype=torch.float64, device=cuda)
            for i in range(steps):
                ops.randn(e, 77, i)
                elbos.append(float(m.elbo_and_grad((xd, yd), epsilon=e).item()))
                opt.step()
        traj.append((elbos, m.q_mu.unconstrained.cpu().numpy().copy(), m.q_sqrt.unconstrained.cpu().numpy().copy(),
                     m.kernel.lengthscales.unconstrained.item()))
    (e0, mu0, sq0, ls0), (e1, mu1, sq1, ls1) = traj
    assert np.allclose(e0, e1, rtol=1e-13, atol=0.0), (e0, e1)
    assert _rel(mu1, mu0) < 1e-13 and _rel(sq1, sq0) < 1e-13 and abs(ls1 - ls0) < 1e-14


def test_fused_q_sqrt_adam_is_bit_identical(cuda):
    """fcest_tri_grad_adam (Adam update of q_sqrt inside tri_grad's epilogue) against tri_grad + fcest_adam_step_dev:
    same trajectory bit for bit; the strict upper triangle of q_sqrt and of its moments is never touched."""
    from fcest_b200.helpers.inference import GraphedTrainingStep, KerasAdam
    N, D, S, M = 60, 6, 2, 40
    out = []
    for fuse in (False, True):
        m, x, y, eps, p = _build("svgp", N, D, S, M)
        m.likelihood.seed(5)
        opt = KerasAdam(m.trainable_parameters)
        g = GraphedTrainingStep(m, (x.numpy(), y.numpy()), opt, fuse_q_sqrt=fuse)
        elbos = [float(g().item()) for _ in range(4)]
        g.check()
        idx = [i for i, q in enumerate(opt.parameters) if q is m.q_sqrt][0]
        out.append((elbos, m.q_sqrt.unconstrained.cpu().numpy().copy(), opt.m[idx].cpu().numpy().copy(),
                    opt.v[idx].cpu().numpy().copy(), m.q_mu.unconstrained.cpu().numpy().copy()))
        assert (m.q_sqrt.grad is None) == fuse
    a, b = out
    assert a[0] == b[0]
    for u, w in zip(a[1:], b[1:]):
        assert np.array_equal(u, w)
    iu = np.triu_indices(M, k=1)
    assert not b[1][:, iu[0], iu[1]].any() and not b[2][:, iu[0], iu[1]].any() and not b[3][:, iu[0], iu[1]].any()


def test_full_size_properties_c4(cuda):
    """BASELINE.json's metric configuration (C4: N=1200, D=50, S=8, M=200), where the oracle is too slow to be the
    checker: size-independent properties instead.
      1. additivity over time: the ELBO and every gradient of the full evaluation equal the sums over two time
         halves with the KL term weighted 1/2 each (the identity the time-sharded multi-GPU mode rests on);
      2. invariance under a permutation of the Monte-Carlo samples (mean over S);
      3. the analytic gradient against central differences of the forward-only ELBO along random directions in
         q_mu, q_sqrt, A and lambda (same noise)."""
    from fcest_b200.models import SparseVariationalWishartProcess
    N, D, S, M = 1200, 50, 8, 200
    x, y, _, p = wo.make_inputs(N, D, 1, M=M, seed=0)
    m = SparseVariationalWishartProcess(D, p["Z"].numpy(), num_mc_samples=S, verbose=False)
    m.q_mu.assign(p["q_mu"]); m.q_sqrt.assign(p["q_sqrt"])
    m.likelihood.A_scale_matrix.assign(p["A"]); m.likelihood.additive_part.assign(p["lam"])
    xd, yd = x.cuda(), y.cuda()
    g = torch.Generator(device="cuda").manual_seed(11)
    eps = torch.randn(S, N, D * D, dtype=torch.float64, device="cuda", generator=g)
    params = [m.q_mu, m.q_sqrt, m.kernel.variance, m.kernel.lengthscales, m.likelihood.A_scale_matrix,
              m.likelihood.additive_part, m.inducing_variable.Z]
    full = m.elbo_and_grad((xd, yd), epsilon=eps).item()
    g_full = [q.grad.clone() for q in params]
    # 1. two time halves
    parts, g_sum = 0.0, [torch.zeros_like(t) for t in g_full]
    for sl in (slice(0, 600), slice(600, N)):
        parts += m.elbo_and_grad((xd[sl], yd[sl]), epsilon=eps[:, sl].contiguous(), kl_scale=0.5).item()
        for acc, q in zip(g_sum, params):
            acc += q.grad
    assert abs(parts - full) <= 1e-11 * abs(full)
    for a, b, q in zip(g_sum, g_full, params):
        if q is m.q_sqrt:
            # tri_grad takes the whole KL term in every evaluation (the time-sharded mode all-reduces the packed
            # cotangent before it), so the two halves count d KL / d q_sqrt = tril(L) - diag(1 / L_mm) twice
            L = torch.tril(m.q_sqrt.unconstrained)
            kl_grad = L - torch.diag_embed(1.0 / torch.diagonal(L, dim1=-2, dim2=-1))
            a = a + kl_grad
        assert _rel(a.cpu().numpy().reshape(-1), b.cpu().numpy().reshape(-1)) < 1e-8, q.name
    # 2. permutation of the samples
    perm = torch.tensor([3, 7, 0, 5, 1, 6, 2, 4], device="cuda")
    e2 = m.elbo((xd, yd), epsilon=eps[perm].contiguous()).item()
    assert abs(e2 - full) <= 1e-11 * abs(full)
    # 3. directional derivatives (central differences of the forward-only ELBO, same noise).  The direction leans on
    #    the gradient itself so that the derivative is large against the rounding of a -4e6 ELBO; q_sqrt moves on
    #    its strict lower triangle only (its diagonal enters through log L_mm^2 with entries near zero here)
    rng = np.random.default_rng(2)
    for q, gq, h in [(m.q_mu, g_full[0], 1e-4), (m.q_sqrt, g_full[1], 1e-5), (m.likelihood.A_scale_matrix, g_full[4], 1e-5),
                     (m.likelihood.additive_part, g_full[5], 1e-4)]:
        gdir = torch.tril(gq, diagonal=-1) if q is m.q_sqrt else gq
        d = gdir / gdir.abs().max() + 0.3 * torch.tensor(rng.standard_normal(tuple(gq.shape)), device="cuda")
        if q is m.q_sqrt:
            d = torch.tril(d, diagonal=-1)
        base = q.unconstrained.clone()
        q.unconstrained.copy_(base + h * d); ep = m.elbo((xd, yd), epsilon=eps).item()
        q.unconstrained.copy_(base - h * d); em = m.elbo((xd, yd), epsilon=eps).item()
        q.unconstrained.copy_(base)
        fd = (ep - em) / (2 * h)
        an = float((gq.reshape(-1) * d.reshape(-1)).sum().item())
        assert abs(fd - an) <= 1e-4 * abs(an), (q.name, fd, an)


def test_elbo_raises_on_failed_cholesky(cuda):
    """model.elbo() inspects the Cholesky status like elbo_and_grad (tf.linalg.cholesky raises in the reference,
    likelihoods.py:184-189): Sigma = 0 when A = 0 and softplus(lambda) underflows."""
    m, x, y, eps, p = _build("svgp", 20, 3, 2, 6)
    m.likelihood.A_scale_matrix.assign(np.zeros((3, 3)))
    m.likelihood.additive_part.assign(np.full(3, -800.0))
    with pytest.raises(RuntimeError):
        m.elbo((x.numpy(), y.numpy()), epsilon=eps.cuda())
    with pytest.raises(RuntimeError):
        m.elbo_and_grad((x.numpy(), y.numpy()), epsilon=eps.cuda())
    assert not np.isfinite(m.elbo((x.numpy(), y.numpy()), epsilon=eps.cuda(), check=False).item())


def test_graph_replay_survives_a_larger_eager_workspace_request(cuda):
    """A captured training step holds raw pointers into the split-K scratch buffers; an eager GEMM that outgrows
    the cached buffer between replays must not free them (ops._workspace retires, never releases)."""
    from fcest_b200 import ops
    from fcest_b200.helpers.inference import GraphedTrainingStep, KerasAdam
    m, x, y, eps, p = _build("svgp", 40, 4, 3, 12)
    m2, *_ = _build("svgp", 40, 4, 3, 12)
    data = (x.numpy(), y.numpy())
    for mm in (m, m2):
        mm.likelihood.seed(9)
    g = GraphedTrainingStep(m, data, KerasAdam(m.trainable_parameters))
    ref = GraphedTrainingStep(m2, data, KerasAdam(m2.trainable_parameters))
    for _ in range(3):
        g(); ref()
    before = dict(ops._ws_cache)
    # outgrow every cached scratch buffer on this device's streams, then release what the allocator can reuse
    for key, buf in before.items():
        s = torch.cuda.ExternalStream(key[1]) if key[1] else torch.cuda.default_stream()
        with torch.cuda.stream(s):
            ops._workspace(buf.numel() * 8 + (4 << 20), buf.device)
    assert all(any(b is r for r in ops._ws_retired) for b in before.values())
    junk = [torch.full((1 << 18,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(16)]
    torch.cuda.synchronize()
    for _ in range(3):
        g(); ref()
    del junk
    assert _rel(m.q_sqrt.unconstrained.cpu().numpy(), m2.q_sqrt.unconstrained.cpu().numpy()) < 1e-13
    assert _rel(m.q_mu.unconstrained.cpu().numpy(), m2.q_mu.unconstrained.cpu().numpy()) < 1e-13


@pytest.mark.parametrize("kind,N,D,S,M", [("svgp", 40, 3, 4, 12), ("vgp", 30, 3, 5, None)])
def test_graphed_evaluation_matches_eager(cuda, kind, N, D, S, M):
    """model.compile_evaluation: the captured ELBO + gradient evaluation is bit-identical to the eager one and sees
    parameter updates made between replays."""
    m, x, y, eps, p = _build(kind, N, D, S, M)
    data = (x.numpy(), y.numpy()) if kind == "svgp" else None
    f = m.compile_evaluation(data)
    e = eps.cuda()
    for it in range(4):
        if it == 3:
            m.q_mu.unconstrained.mul_(1.5)           # in place: the graph reads the live parameter tensors
        got = f(epsilon=e).item()
        g_q = m.q_sqrt.grad.clone(); g_A = m.likelihood.A_scale_matrix.grad.clone(); g_mu = m.q_mu.grad.clone()
        ref = m.elbo_and_grad(data, epsilon=e).item()
        assert got == ref, (it, got, ref)
        assert torch.equal(g_q, m.q_sqrt.grad) and torch.equal(g_A, m.likelihood.A_scale_matrix.grad)
        assert torch.equal(g_mu, m.q_mu.grad)
    f.check()
    elbo_ref, _ = wo.elbo_and_grad(kind, {**p, "q_mu": p["q_mu"] * 1.5}, x, y, eps)
    assert abs(got - elbo_ref) <= 1e-8 * abs(elbo_ref)
