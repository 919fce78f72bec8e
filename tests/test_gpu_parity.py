"""Parity of the CUDA path (called through the C ABI) against the CPU oracle -- GPU box only.

Tolerances.  The north-star bar is 1e-10 relative in the 2-norm for the PC apply.  The
frequency-domain systems have condition number ~ 12 sqrt(gamma) / h^2, so two correct
fp64 algorithms may differ by cond * eps: at N_x <= 1024 that stays below 1e-10 and the bar is
asserted as is; at larger N_x it is asserted where it is reachable and otherwise replaced by
"no further from the 80-bit truth than the fp64 oracle itself (x3)".
"""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from optimal_control_paradiag_b200 import DiagFFTPC, ParaDiagHandle, petsc_shim  # noqa: E402
from oracle.gmres import gmres as oracle_gmres  # noqa: E402
from oracle.operator import AllAtOnce  # noqa: E402
from oracle.pc_fast import DiagFFTPCFast  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
PC_TOL = 1e-10


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def rand_x(size, seed=0, real=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(size) + 0j
    if not real:
        x = x + 1j * rng.standard_normal(size)
    return x


# ---------------------------------------------------------------------------- time-axis FFT
@pytest.mark.parametrize("variant", ["l2", "cluster", "tma"])
def test_fft_16384_both_kernels_many_lines(variant, monkeypatch):
    # both N_t = 16384 kernels (cluster-free default, 4-CTA cluster), enough lines for every persistent CTA to
    # loop: out of place and in place.  (A remote store overtaking a not-yet-performed shared-memory load in
    # the cluster kernel corrupted about one line in 5000 before the block-scope fence went in.)
    import scipy.fft as sfft
    monkeypatch.setenv("PD_FFT16K", variant)
    N_t, nl = 16384, 1500
    perm = np.concatenate([np.arange(q, N_t, 4) for q in range(4)])
    x = rand_x(nl * N_t).reshape(nl, N_t)
    with ParaDiagHandle(8, N_t) as h:
        xt = torch.tensor(x, device=DEV).reshape(-1)
        yt, zt = torch.empty_like(xt), torch.empty_like(xt)
        for rep in range(3):
            h.stage_fft(xt, yt, nl, True)
            if rep == 0:
                assert rel(yt.cpu().numpy().reshape(nl, N_t), sfft.ifft(x, axis=1)[:, perm]) < 5e-15
            h.stage_fft(yt, zt, nl, False)
            assert np.abs(zt.cpu().numpy().reshape(nl, N_t) - x).max() < 1e-12
            h.stage_fft(yt, yt, nl, False)
            assert np.abs(yt.cpu().numpy().reshape(nl, N_t) - x).max() < 1e-12
            zt.copy_(xt)
            h.stage_fft(zt, zt, nl, True)
            h.stage_fft(zt, zt, nl, False)
            assert np.abs(zt.cpu().numpy().reshape(nl, N_t) - x).max() < 1e-12


@pytest.mark.parametrize("N_t", [3, 5, 13, 64, 81, 96, 97, 100, 128, 256, 512, 625, 1024, 2048, 4096, 8192, 16384])
def test_fft_matches_scipy(N_t):
    import scipy.fft as sfft
    nl = 19
    x = rand_x(nl * N_t).reshape(nl, N_t)
    with ParaDiagHandle(8, N_t) as h:
        xt = torch.tensor(x, device=DEV).reshape(-1)
        yt = torch.empty_like(xt)
        if N_t == 16384:
            # N_t = 16384 kernels: time -> frequency leaves [k = 0 mod 4 | 1 | 2 | 3]; frequency -> time
            # consumes that order
            perm = np.concatenate([np.arange(q, N_t, 4) for q in range(4)])
            h.stage_fft(xt, yt, nl, True)
            assert rel(yt.cpu().numpy().reshape(nl, N_t), sfft.ifft(x, axis=1)[:, perm]) < 5e-15
            xp = torch.tensor(np.ascontiguousarray(x[:, perm]), device=DEV).reshape(-1)
            h.stage_fft(xp, yt, nl, False)
            assert rel(yt.cpu().numpy().reshape(nl, N_t), sfft.fft(x, axis=1)) < 5e-15
            h.stage_fft(xt, yt, nl, True)
            h.stage_fft(yt, yt, nl, False)
            assert rel(yt.cpu().numpy().reshape(nl, N_t), x) < 5e-15
            return
        h.stage_fft(xt, yt, nl, False)
        assert rel(yt.cpu().numpy().reshape(nl, N_t), sfft.fft(x, axis=1)) < 5e-15
        h.stage_fft(xt, yt, nl, True)
        assert rel(yt.cpu().numpy().reshape(nl, N_t), sfft.ifft(x, axis=1)) < 5e-15
        h.stage_fft(yt, yt, nl, False)                      # in place, round trip
        assert rel(yt.cpu().numpy().reshape(nl, N_t), x) < 5e-15


# ------------------------------------------------------------------------------- PC apply
SMALL = [(20, 16384, 1.0), (2, 3, 1.0), (3, 4, 1.0), (16, 13, 1.0), (17, 64, 1.0), (18, 64, 1.0), (34, 8, 1.0), (35, 12, 1e-2),
         (16, 16, 1.0), (20, 81, 1.0), (80, 81, 1.0), (24, 64, 1e-4), (40, 96, 1e-2), (100, 128, 1.0),
         (257, 60, 1.0), (256, 256, 1.0), (300, 81, 1e-6)]


@pytest.mark.parametrize("N_x,N_t,gamma", SMALL)
def test_pc_apply_matches_oracle(N_x, N_t, gamma):
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        x = rand_x(h.size)
        ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(x)
        y = h.pc_apply(torch.tensor(x, device=DEV)).cpu().numpy()
        assert rel(y, ref) < PC_TOL
        # Dirichlet rows: exactly zero whatever the input holds there
        assert np.abs(y.reshape(2, N_x + 1, N_t)[:, [0, -1], :]).max() == 0.0
        # host-buffer entry point and in-place device apply give the same bits
        assert np.array_equal(h.pc_apply_host(x), y)
        xt = torch.tensor(x, device=DEV)
        h.pc_apply(xt, xt)
        assert np.array_equal(xt.cpu().numpy(), y)
        # real input (what GMRES actually feeds the PC) -> real output
        xr = rand_x(h.size, seed=1, real=True)
        yr = h.pc_apply(torch.tensor(xr, device=DEV)).cpu().numpy()
        assert rel(yr, DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(xr)) < PC_TOL
        assert np.abs(yr.imag).max() <= 1e-11 * np.abs(yr.real).max()


@pytest.mark.parametrize("N_x,N_t,gamma", [(16, 128, 1.0), (33, 256, 1e-2), (100, 512, 1.0), (64, 1024, 1e-4),
                                           (20, 2048, 1.0), (17, 4096, 1.0), (9, 8192, 1.0), (8, 16384, 1.0),
                                           (1024, 1024, 1.0),
                                           # the shared-memory pair kernel: odd, prime, composite and small
                                           # power-of-two N_t (the upstream default N_t = 81 first)
                                           (80, 81, 1.0), (16, 8, 1.0), (16, 9, 1.0), (33, 13, 1e-2), (20, 16, 1.0),
                                           (64, 64, 1.0), (40, 100, 1.0), (300, 96, 1e-4), (50, 127, 1.0),
                                           (12, 1000, 1.0), (1024, 243, 1.0)])
def test_real_input_fast_path_matches_oracle(N_x, N_t, gamma):
    # pd_pc_apply_real: half spectrum (k <= N_t/2), real vectors in and out
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        x = np.random.default_rng(1).standard_normal(h.size)
        ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(x + 0j)
        xt = torch.tensor(x, device=DEV)
        y = h.pc_apply_real(xt).cpu().numpy()
        assert float(np.linalg.norm(y - ref.real) / np.linalg.norm(ref)) < PC_TOL
        yc = h.pc_apply(torch.tensor(x + 0j, device=DEV)).cpu().numpy()
        # register pipelines: same kernels, half the columns; the shared-memory pair kernel orders its sums differently
        same = N_t >= 128 and N_t & (N_t - 1) == 0
        assert float(np.linalg.norm(y - yc.real) / np.linalg.norm(yc)) < (1e-13 if same else 1e-11)
        assert np.abs(y.reshape(2, N_x + 1, N_t)[:, [0, -1], :]).max() == 0.0
        h.pc_apply_real(xt, xt)                                                      # in place
        assert np.array_equal(xt.cpu().numpy(), y)


def test_real_input_fast_path_unsupported_sizes_fail_loudly():
    from optimal_control_paradiag_b200 import ParaDiagError
    with ParaDiagHandle(16, 7) as h:        # a half-spectrum row (8 columns) would not fit the 7-column workspaces
        assert not h.real_path_supported
        x = torch.zeros(h.size, dtype=torch.float64, device=DEV)
        with pytest.raises(ParaDiagError):
            h.pc_apply_real(x)
        with pytest.raises(ParaDiagError):
            h.gmres_real(x)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "pc_apply_*.npz"))))
def test_pc_apply_matches_golden(path):
    g = np.load(path)
    with ParaDiagHandle(int(g["N_x"]), int(g["N_t"]), T=float(g["T"]), gamma=float(g["gamma"])) as h:
        assert rel(h.pc_apply_host(g["x"]), g["y"]) < PC_TOL
        assert rel(h.pc_apply_host(g["x_real"]), g["y_real"]) < PC_TOL


def test_pc_apply_config2_1024x1024():
    N_x = N_t = 1024
    with ParaDiagHandle(N_x, N_t) as h:
        x = rand_x(h.size)
        ref = DiagFFTPCFast(N_x, N_t).apply(x)
        y = h.pc_apply(torch.tensor(x, device=DEV)).cpu().numpy()
        assert rel(y, ref) < PC_TOL


def test_pc_apply_gamma_sweep_4096_small_gamma():
    # config 5 family (N_x = 4096); N_t kept small so the CPU oracle finishes in seconds
    N_x, N_t = 4096, 64
    for gamma in (1e-4, 1e-6):
        with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
            x = rand_x(h.size)
            ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(x)
            assert rel(h.pc_apply_host(x), ref) < PC_TOL


@pytest.mark.parametrize("N_x,N_t", [(9536, 8), (9600, 8), (16384, 12), (20000, 5), (65536, 4)])
def test_pc_apply_deep_partition_levels_against_long_double(N_x, N_t):
    # 3 and 4 partition levels (rows 9599 -> 564 -> 33 -> 1 etc.); compared with the 80-bit oracle
    with ParaDiagHandle(N_x, N_t) as h:
        x = rand_x(h.size)
        truth = DiagFFTPCFast(N_x, N_t, dtype=np.longdouble).apply(x)
        y = h.pc_apply_host(x)
        assert float(np.linalg.norm(y - truth) / np.linalg.norm(truth)) < 1e-10


def test_cuda_path_is_closer_to_the_truth_than_fp64_lu():
    # detuning-form coefficients: no cancellation near the discrete wave resonances (pd_solve.cu)
    N_x, N_t = 4096, 128
    with ParaDiagHandle(N_x, N_t) as h:
        for x in (rand_x(h.size), AllAtOnce(N_x, N_t).rhs() + 0j):
            truth = DiagFFTPCFast(N_x, N_t, dtype=np.longdouble).apply(x)
            e_oracle = float(np.linalg.norm(DiagFFTPCFast(N_x, N_t).apply(x) - truth) / np.linalg.norm(truth))
            e_cuda = float(np.linalg.norm(h.pc_apply_host(x) - truth) / np.linalg.norm(truth))
            assert e_cuda < 1e-10 and e_cuda < 0.2 * e_oracle, (e_cuda, e_oracle)


def test_pc_apply_conditioning_limited_case_against_long_double():
    # gamma = 1, N_x = 4096: cond ~ 2e8, two fp64 algorithms differ by ~1e-9; compare both with 80-bit
    N_x, N_t = 4096, 32
    with ParaDiagHandle(N_x, N_t) as h:
        x = rand_x(h.size)
        truth = DiagFFTPCFast(N_x, N_t, dtype=np.longdouble).apply(x)
        f64 = DiagFFTPCFast(N_x, N_t).apply(x)
        y = h.pc_apply_host(x)
        e_oracle = float(np.linalg.norm(f64 - truth) / np.linalg.norm(truth))
        e_cuda = float(np.linalg.norm(y - truth) / np.linalg.norm(truth))
        assert e_cuda < max(3 * e_oracle, PC_TOL), (e_cuda, e_oracle)
        assert rel(y, f64) < 1e-7


# ------------------------------------------- full BASELINE sizes: size-independent properties
def _properties(N_x, N_t, gamma=1.0):
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        g = torch.Generator(device=DEV).manual_seed(0)
        x = torch.randn(h.size, dtype=torch.float64, device=DEV, generator=g) \
            + 1j * torch.randn(h.size, dtype=torch.float64, device=DEV, generator=g)
        y = h.pc_apply(x)
        X = x.view(2, N_x + 1, N_t)
        # (1) P (P^-1 x) = x on interior rows, P the explicit block-circulant stencil
        # measured as the normwise backward error ||P y - x|| / (||P|| ||y||), ||P|| bounded by
        # ||C1|| ||M|| + dt^2/2 ||C2|| ||K|| + c ||M|| <= 4h + 4 dt^2/h + c h
        r = h.pc_matvec(y).view(2, N_x + 1, N_t)[:, 1:-1, :] - X[:, 1:-1, :]
        hh, dt = 1.0 / N_x, 2.0 / N_t
        normP = 4 * hh + 4 * dt * dt / hh + dt * dt / gamma ** 0.5 * hh
        back = float(torch.linalg.norm(r) / (normP * torch.linalg.norm(y)))
        res = float(torch.linalg.norm(r) / torch.linalg.norm(X[:, 1:-1, :]))
        # (2) boundary rows exactly zero
        Y = y.view(2, N_x + 1, N_t)
        bnd = float(Y[:, [0, -1], :].abs().max())
        # (3) linearity
        z = torch.randn(h.size, dtype=torch.float64, device=DEV, generator=g).to(torch.complex128)
        a, b = 0.75 - 0.5j, -1.25 + 2.0j
        lin = h.pc_apply(a * x + b * z) - (a * y + b * h.pc_apply(z))
        lin = float(torch.linalg.norm(lin) / torch.linalg.norm(y))
        # (4) real in -> real out
        yz = h.pc_apply(z)
        im = float(yz.imag.abs().max() / yz.real.abs().max())
        return back, res, bnd, lin, im


@pytest.mark.parametrize("N_x,N_t", [(1024, 1024), (4096, 4096), (16384, 4096)])
def test_pc_apply_properties_at_baseline_sizes(N_x, N_t):
    back, res, bnd, lin, im = _properties(N_x, N_t)
    assert back < 1e-14, back          # backward error of the inverse: not conditioning-limited
    assert res < 1e-7, res             # plain residual ||P y - x|| / ||x|| (grows with ||y||/||x||)
    assert bnd == 0.0
    assert lin < 1e-8, lin             # forward error of differences is cond-limited
    assert im < 1e-8, im


# ------------------------------------------------------------------ operator, rhs, GMRES
@pytest.mark.parametrize("N_x,N_t,gamma", [(16, 13, 1.0), (80, 81, 1.0), (32, 64, 1e-2), (64, 256, 1e-4), (257, 100, 0.5)])
def test_matvec_and_rhs_match_oracle(N_x, N_t, gamma):
    op = AllAtOnce(N_x, N_t, 2.0, gamma)
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        x = rand_x(h.size, seed=1)
        assert rel(h.matvec(torch.tensor(x, device=DEV)).cpu().numpy(), op.matvec(x)) < 1e-14
        assert rel(h.build_rhs().cpu().numpy(), op.rhs()) < 1e-13
    with ParaDiagHandle(N_x, N_t, gamma=gamma, bug138=False) as h:
        op2 = AllAtOnce(N_x, N_t, 2.0, gamma, bug138=False)
        assert rel(h.matvec(torch.tensor(x, device=DEV)).cpu().numpy(), op2.matvec(x)) < 1e-14


def test_pc_matvec_is_the_explicit_circulant_matrix():
    from oracle.pc_explicit import ExplicitPC
    N_x, N_t, gamma = 12, 10, 0.3
    e = ExplicitPC(N_x, N_t, 2.0, gamma)
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        x = rand_x(h.size)
        y = h.pc_matvec(torch.tensor(x, device=DEV)).cpu().numpy().reshape(2, N_x + 1, N_t)
        want = (e.P @ x.reshape(2, N_x + 1, N_t)[:, 1:-1, :].reshape(-1)).reshape(2, N_x - 1, N_t)
        assert rel(y[:, 1:-1, :], want) < 1e-14


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "gmres_*.npz"))))
def test_gmres_matches_golden(path):
    g = np.load(path)
    N_x, N_t, gamma = int(g["N_x"]), int(g["N_t"]), float(g["gamma"])
    with ParaDiagHandle(N_x, N_t, T=float(g["T"]), gamma=gamma) as h:
        b = h.build_rhs()
        assert rel(b.cpu().numpy(), g["b"]) < 1e-13
        x, its, hist, reason = h.gmres(b, rtol=1e-7)
        assert reason == "CONVERGED_RTOL"
        assert abs(its - int(g["its"])) <= 1                      # north star: +-1
        assert np.allclose(hist[: its], g["hist"][: its], rtol=1e-5)
        assert rel(x.cpu().numpy(), g["x"]) < 1e-7
        # random right-hand side: O(50-100) iterations, still within +-1
        rng = np.random.default_rng(0)
        br = rng.standard_normal((2, N_x + 1, N_t))
        br[:, 0] = br[:, -1] = 0
        _, its_r, hist_r, _ = h.gmres(torch.tensor(br.reshape(-1) + 0j, device=DEV), rtol=1e-7)
        assert abs(its_r - int(g["its_random"])) <= 1


def test_gmres_restart_and_max_it():
    N_x, N_t = 16, 24
    op = AllAtOnce(N_x, N_t)
    pc = DiagFFTPCFast(N_x, N_t)
    b = np.random.default_rng(0).standard_normal(2 * (N_x + 1) * N_t) + 0j
    _, its_o, _, _ = oracle_gmres(op.matvec, pc.apply, b, rtol=1e-8, restart=10, max_it=2000)
    with ParaDiagHandle(N_x, N_t) as h:
        bt = torch.tensor(b, device=DEV)
        x, its, hist, reason = h.gmres(bt, rtol=1e-8, restart=10, max_it=2000)
        assert reason == "CONVERGED_RTOL" and abs(its - its_o) <= 2
        r = pc.apply(op.matvec(x.cpu().numpy()) - b)
        assert np.linalg.norm(r) <= 1.05e-8 * np.linalg.norm(pc.apply(b))
        x, its, hist, reason = h.gmres(bt, rtol=1e-14, restart=300, max_it=7)
        assert its == 7 and reason == "DIVERGED_ITS" and len(hist) == 8


def test_gmres_iteration_parity_at_4096():
    # beyond N_x ~ 2000 the 5-step termination of the manufactured problem is broken by rounding
    # (cond ~ 2e8): the count then depends on how much noise each implementation injects.  The device
    # path (cancellation-free solve and matvec) injects less than the fp64 oracle, so it may need FEWER
    # iterations; it must never need more than the oracle's count + 1 (north star: +-1).
    N_x, N_t = 4096, 512
    from oracle import csolve
    op = AllAtOnce(N_x, N_t)
    pc = DiagFFTPCFast(N_x, N_t, solver=csolve.thomas_toeplitz_c)
    with ParaDiagHandle(N_x, N_t) as h:
        b = h.build_rhs()
        for rtol in (1e-5, 1e-7):
            _, its_o, hist_o, _ = oracle_gmres(op.matvec, pc.apply, op.rhs() + 0j, rtol=rtol, max_it=60)
            _, its, hist, reason = h.gmres(b, rtol=rtol, max_it=60)
            assert reason == "CONVERGED_RTOL" and its <= its_o + 1 and its >= 5, (rtol, its, its_o, hist, hist_o)
            assert np.allclose(hist[:5], hist_o[:5], rtol=1e-6)          # identical until rounding takes over


@pytest.mark.parametrize("N_x,N_t,gamma", [(64, 128, 1.0), (200, 256, 1e-2), (1024, 1024, 1.0), (80, 81, 1.0),
                                           (100, 48, 1.0)])
def test_real_vector_gmres_equals_complex_gmres(N_x, N_t, gamma):
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        bc = h.build_rhs()
        br = h.build_rhs_real()
        assert np.array_equal(br.cpu().numpy(), bc.cpu().numpy().real)
        v = torch.randn(h.size, dtype=torch.float64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
        assert rel(h.matvec_real(v).cpu().numpy(), h.matvec(v.to(torch.complex128)).cpu().numpy().real) < 1e-15
        xc, its_c, hist_c, reason_c = h.gmres(bc, rtol=1e-7)
        xr, its_r, hist_r, reason_r = h.gmres_real(br, rtol=1e-7)
        assert reason_r == "CONVERGED_RTOL" and abs(its_r - its_c) <= 1
        assert np.allclose(hist_r[: min(its_r, its_c)], hist_c[: min(its_r, its_c)], rtol=1e-5)
        assert rel(xr.cpu().numpy(), xc.cpu().numpy().real) < 1e-6
        # and again complex after real on the same handle (basis cache is re-sized)
        xc2, its_c2, _, _ = h.gmres(bc, rtol=1e-7)
        assert its_c2 == its_c
        # random right-hand side, tens of iterations
        rng = np.random.default_rng(0)
        b = rng.standard_normal((2, N_x + 1, N_t))
        b[:, 0] = b[:, -1] = 0
        _, its_rc, _, _ = h.gmres(torch.tensor(b.reshape(-1) + 0j, device=DEV), rtol=1e-7, max_it=400)
        _, its_rr, _, _ = h.gmres_real(torch.tensor(b.reshape(-1), device=DEV), rtol=1e-7, max_it=400)
        assert abs(its_rc - its_rr) <= 1


def test_gmres_config2_manufactured_rhs():
    with ParaDiagHandle(1024, 1024) as h:
        b = h.build_rhs()
        x, its, hist, reason = h.gmres(b, rtol=1e-7)
        assert reason == "CONVERGED_RTOL" and its in (5, 6, 7)
        res = h.matvec(x) - b
        assert float(torch.linalg.norm(res) / torch.linalg.norm(b)) < 1e-4


# ----------------------------------------------------- the reference-facing python-PC surface
def test_diagfftpc_through_petsc_protocol():
    N_x, N_t, gamma = 80, 81, 1.0                      # the script's constants, :335-339
    DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=gamma)
    pc = petsc_shim.PC()
    pc.setPythonType("optimal_control_paradiag_b200.DiagFFTPC")
    pc.setUp()
    x = rand_x(2 * (N_x + 1) * N_t)
    xv, yv = petsc_shim.Vec(x), petsc_shim.Vec.zeros(x.size)
    pc.apply(xv, yv)
    assert rel(yv.getArray(), DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(x)) < PC_TOL
    with pytest.raises(NotImplementedError):
        pc.applyTranspose(xv, yv)
    # torch device tensors take the zero-copy path
    xt = torch.tensor(x, device=DEV)
    yt = torch.empty_like(xt)
    pc.apply(xt, yt)
    assert np.array_equal(yt.cpu().numpy(), yv.getArray())
    pc.destroy()
    DiagFFTPC._defaults = {}


@pytest.mark.parametrize("N_x,N_t", [(80, 81), (64, 256), (33, 1024)])
def test_real_host_vectors_take_the_half_spectrum_path(N_x, N_t):
    # float64 host Vecs (a real-scalar PETSc build / numpy float64): pd_pc_apply_real_host, half the PCIe bytes
    x = np.random.default_rng(5).standard_normal(2 * (N_x + 1) * N_t)
    ref = DiagFFTPCFast(N_x, N_t, 2.0, 1.0).apply(x + 0j).real
    with ParaDiagHandle(N_x, N_t) as h:
        y = h.pc_apply_real_host(x)
        assert y.dtype == np.float64 and rel(y, ref) < PC_TOL
        assert np.array_equal(y, h.pc_apply_real(torch.tensor(x, device=DEV)).cpu().numpy())
        xin = x.copy()
        h.pc_apply_real_host(xin, xin)                               # in place
        assert np.array_equal(xin, y)
    DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=1.0)
    try:
        pc = petsc_shim.PC()
        pc.setPythonContext(DiagFFTPC())
        pc.setUp()
        yv = np.zeros_like(x)
        pc.apply(x, yv)                                              # numpy float64 in, float64 out
        assert np.array_equal(yv, y)
        pc.destroy()
    finally:
        DiagFFTPC._defaults = {}


def test_problem_class_reproduces_the_reference_run():
    from optimal_control_paradiag_b200 import Optimal_Control_Wave_Equation, default_parameters
    equ = Optimal_Control_Wave_Equation(80, 2, 81, 1)
    u_sol, p_sol = equ.solve(parameters=default_parameters, complex=True, verbose=False)
    assert equ.ksp_its == 5 and equ.ksp_reason == "CONVERGED_RTOL"
    assert tuple(u_sol.shape) == (81, 81)
    assert equ.error_norm(u_sol) < 0.1
    with pytest.raises(NotImplementedError):
        equ.solve(parameters={'ksp_type': 'cg', 'pc_type': 'jacobi'}, complex=True)


@pytest.mark.parametrize("N,gamma", [(10, 1.0), (24, 1.0), (30, 1e-2), (40, 1.0)])
def test_direct_lu_baseline_equals_gmres_with_the_pc(N, gamma):
    # the pc=False branch of the upstream script (:186, :573-577) on the product path, against GMRES + DiagFFTPC
    from optimal_control_paradiag_b200 import Optimal_Control_Wave_Equation, default_parameters
    equ = Optimal_Control_Wave_Equation(N, 2, N, gamma)
    u_d, p_d = equ.solve(parameters=None, complex=True, verbose=False)
    u_d, p_d = u_d.clone(), p_d.clone()
    u_g, p_g = equ.solve(parameters=default_parameters, complex=True, verbose=False, rtol=1e-13)
    assert equ.ksp_reason.startswith("CONVERGED")
    err = float(torch.linalg.norm(torch.cat([u_d - u_g, p_d - p_g])) / torch.linalg.norm(torch.cat([u_g, p_g])))
    assert err < 1e-9, err
    # and against the oracle's SuperLU restatement of the same branch
    direct = AllAtOnce(N, N, 2.0, gamma).direct_solve().reshape(2, N + 1, N)
    assert rel(u_d.cpu().numpy().real, direct[0]) < 1e-9
    equ.handle.close()


def test_errors_are_status_codes_not_aborts():
    from optimal_control_paradiag_b200 import ParaDiagError
    with pytest.raises(ParaDiagError):
        ParaDiagHandle(16, 16, device=99)
    with ParaDiagHandle(16, 16) as h:
        x = torch.zeros(h.size, dtype=torch.complex128, device=DEV)
        with pytest.raises(ValueError):
            h.pc_apply(x[:-1], x[:-1])
        with pytest.raises(ParaDiagError):
            h.matvec(x, x)                                # aliasing is rejected
        assert h.launch_count >= 0 and h.workspace_bytes > 0


def test_gmres_solution_equals_the_direct_lu_baseline():
    # the reference's pc=False branch (:573-577, direct MUMPS) restated by the oracle, against the device solve
    N_x, N_t = 24, 32
    direct = AllAtOnce(N_x, N_t).direct_solve()
    with ParaDiagHandle(N_x, N_t) as h:
        x, its, hist, reason = h.gmres(h.build_rhs(), rtol=1e-12)
        assert reason == "CONVERGED_RTOL"
        assert rel(x.cpu().numpy(), direct + 0j) < 1e-9


@pytest.mark.parametrize("name", ["refsetup_81_1.npz", "refsetup_64_0.0001.npz", "refsetup_5_1.npz"])
def test_pc_apply_against_the_route_driven_by_executed_upstream_setup(name):
    # S, S^-1, Sigma, Lambda_2 as computed by the EXECUTED upstream lines :387-436 (tests/golden/
    # make_reference_setup_golden.py) drive the line-by-line route; the CUDA path regenerates all of it in-kernel
    from oracle.pc_ref_route import DiagFFTPCRefRoute
    g = np.load(os.path.join(GOLDEN, name))
    N_t, T, gamma = int(g["N_t"]), float(g["T"]), float(g["gamma"])
    N_x = 40
    pc = DiagFFTPCRefRoute(N_x, N_t, T, gamma)
    assert np.array_equal(pc.Sigma, np.stack([g["Sigma_1"], g["Sigma_2"]], -1))
    pc.Lambda_1, pc.Lambda_2 = g["Lambda_1"], g["Lambda_2"]
    pc.S = np.stack([np.stack([g["S11"], g["S12"]], -1), np.stack([g["S21"], g["S22"]], -1)], -2)
    pc.SI = np.stack([np.stack([g["SI11"], g["SI12"]], -1), np.stack([g["SI21"], g["SI22"]], -1)], -2)
    with ParaDiagHandle(N_x, N_t, T=T, gamma=gamma) as h:
        x = rand_x(h.size)
        y = h.pc_apply(torch.tensor(x, device=DEV)).cpu().numpy()
        assert rel(y, pc.apply(x)) < PC_TOL


# ------------------------------------------------ GMRES iteration counts at BASELINE sizes, two-sided
def _golden_counts():
    import json
    return json.load(open(os.path.join(GOLDEN, "gmres_counts.json")))


@pytest.mark.parametrize("case", ["cfg1", "cfg2", "mid", "cfg5"])
def test_gmres_iteration_counts_two_sided_against_the_oracle(case):
    """North star: "GMRES iteration counts match within +-1".  tests/golden/gmres_counts.json holds the oracle's
    counts (rtol 1e-5 / 1e-7) and how they move when the PC output is perturbed by 1e-13 / 1e-11 relative -- a
    stand-in for another, equally valid fp64 implementation of the same preconditioner.

    * where the oracle's count does NOT move under the perturbation it is a property of the problem: asserted
      two-sided, |its_gpu - its_oracle| <= 1;
    * where it moves (N_x >= 4096: the 5-step termination of the manufactured problem is destroyed by rounding, the
      count then measures the noise each implementation injects) a +-1 comparison between ANY two implementations
      is meaningless; asserted: at least the exact-arithmetic 5 steps, at most the largest count the oracle shows
      under perturbation + 1, and the first five residuals identical to 1e-6.
    """
    g = _golden_counts()[case]
    N_x, N_t, gamma = int(g["N_x"]), int(g["N_t"]), float(g["gamma"])
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        b = h.build_rhs()
        for rtol in ("1e-05", "1e-07"):
            base = g["its"][rtol]
            variants = [base] + [g["perturbed"][e][rtol] for e in g["perturbed"]]
            _, its, hist, reason = h.gmres(b, rtol=float(rtol), max_it=80)
            assert reason == "CONVERGED_RTOL", (case, rtol, its, hist)
            if len(set(variants)) == 1:
                assert abs(its - base) <= 1, (case, rtol, its, variants)
            else:
                assert 5 <= its <= max(v for v in variants if v is not None) + 1, (case, rtol, its, variants)
            k = min(5, its)
            assert np.allclose(hist[:k], g["hist"][:k], rtol=1e-6), (hist[:k], g["hist"][:k])
            # the float64 solve (half-spectrum PC) on the same problem: within the same window
            if h.real_path_supported:
                _, its_r, _, reason_r = h.gmres_real(h.build_rhs_real(), rtol=float(rtol), max_it=80)
                assert reason_r == "CONVERGED_RTOL"
                if len(set(variants)) == 1:
                    assert abs(its_r - base) <= 1, (case, rtol, its_r, variants)
                else:
                    assert 5 <= its_r <= max(v for v in variants if v is not None) + 1, (case, rtol, its_r, variants)


@pytest.mark.parametrize("N_x,N_t,gamma", [(16, 13, 1.0), (9, 3, 1.0), (10, 4, 0.5), (80, 81, 1.0), (257, 100, 0.25),
                                           (64, 256, 1e-4)])
def test_delta_is_A_minus_P(N_x, N_t, gamma):
    # the residual-correction operator (pd_delta) against the oracle's term-by-term (A - P) x and against the
    # device's own matvec - pc_matvec
    op = AllAtOnce(N_x, N_t, 2.0, gamma)
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        x = rand_x(h.size, seed=4).reshape(2, N_x + 1, N_t)
        x[:, 0] = x[:, -1] = 0
        x = x.reshape(-1)
        xt = torch.tensor(x, device=DEV)
        d = h.delta(xt).cpu().numpy()
        want = op.delta(x)
        scale = np.abs(op.matvec(x)).max()
        assert np.abs(d - want).max() < 1e-14 * scale
        dev = (h.matvec(xt) - h.pc_matvec(xt)).cpu().numpy()
        assert np.abs(d - dev).max() < 1e-13 * scale
        # only <= 3 time levels per field are touched
        nz = np.abs(d.reshape(2, N_x + 1, N_t)).max(axis=1) > 0
        assert nz[0].sum() <= 3 and nz[1].sum() <= 2
        # float64 vectors
        xr = torch.tensor(x.real.copy(), device=DEV)
        assert np.abs(h.delta(xr).cpu().numpy() - op.delta(x.real)).max() < 1e-14 * scale


@pytest.mark.parametrize("N_x,N_t,gamma", [(80, 81, 1.0), (32, 64, 1e-2), (1024, 1024, 1.0), (4096, 512, 1.0)])
def test_gmres_residual_correction_mode(N_x, N_t, gamma):
    # same Krylov solve with the preconditioned operator formed as v + P^-1 (A - P) v: same solution, and never
    # more iterations than the default order of operations
    with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
        b = h.build_rhs()
        x0, its0, hist0, r0 = h.gmres(b, rtol=1e-7, max_it=80)
        x1, its1, hist1, r1 = h.gmres(b, rtol=1e-7, max_it=80, correction=True)
        assert r0 == r1 == "CONVERGED_RTOL"
        assert 5 <= its1 <= its0 + 1, (its0, its1)
        assert np.allclose(hist1[:5], hist0[:5], rtol=1e-6)
        res = lambda x: float(torch.linalg.norm(h.matvec(x) - b) / torch.linalg.norm(b))
        assert res(x1) < max(3 * res(x0), 1e-6), (res(x0), res(x1))
        if h.real_path_supported:
            xr, its_r, _, rr = h.gmres_real(h.build_rhs_real(), rtol=1e-7, max_it=80, correction=True)
            assert rr == "CONVERGED_RTOL" and abs(its_r - its1) <= 1
            assert rel(xr.cpu().numpy(), x1.cpu().numpy().real) < 1e-5


# ------------------------------------------- the two interface solvers give the same answer
@pytest.mark.parametrize("N_x,N_t,gamma", [(600, 64, 1.0), (1024, 1024, 1.0), (4096, 128, 1.0), (9600, 8, 1.0),
                                           (16384, 12, 1e-2), (65536, 4, 1.0)])
def test_multilevel_and_sequential_interface_solvers_agree(N_x, N_t, gamma, monkeypatch):
    # default: the one-launch sequential interface kernel (plan-time pivots, cp.async ring); PD_ITHOMAS_MAX=0: the
    # recursive reduce / PCR / back chain.  Both against the 80-bit oracle, and against each other.
    x = rand_x(2 * (N_x + 1) * N_t)
    truth = DiagFFTPCFast(N_x, N_t, 2.0, gamma, dtype=np.longdouble).apply(x)
    ys = []
    for mode in ("8192", "0"):
        monkeypatch.setenv("PD_ITHOMAS_MAX", mode)
        with ParaDiagHandle(N_x, N_t, gamma=gamma) as h:
            y = h.pc_apply_host(x)
            assert float(np.linalg.norm(y - truth) / np.linalg.norm(truth)) < 1e-10, mode
            ys.append(y)
    assert rel(ys[0], ys[1]) < 1e-11
