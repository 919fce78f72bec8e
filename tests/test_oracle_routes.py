"""The oracle's three independent routes to P^-1 x agree, and match the committed fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle.pc_explicit import ExplicitPC
from oracle.pc_fast import DiagFFTPCFast, thomas_toeplitz
from oracle.pc_ref_route import DiagFFTPCRefRoute

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [(16, 13, 1.0), (20, 81, 1.0), (16, 13, 0.01), (16, 16, 1.0), (24, 64, 1e-4), (12, 24, 1.0)]


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def rand_x(N_x, N_t, seed=0):
    rng = np.random.default_rng(seed)
    size = 2 * (N_x + 1) * N_t
    return rng.standard_normal(size) + 1j * rng.standard_normal(size)


@pytest.mark.parametrize("N_x,N_t,gamma", CASES)
def test_three_routes_agree(N_x, N_t, gamma):
    x = rand_x(N_x, N_t)
    e = ExplicitPC(N_x, N_t, 2.0, gamma).apply(x)
    r = DiagFFTPCRefRoute(N_x, N_t, 2.0, gamma).apply(x)
    f = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(x)
    assert rel(r, e) < 2e-12
    assert rel(f, e) < 2e-12
    assert rel(f, r) < 2e-12


def test_riesz_roundtrip_is_algebraically_void():
    x = rand_x(20, 27)
    a = DiagFFTPCRefRoute(20, 27, riesz_roundtrip=True).apply(x)
    b = DiagFFTPCRefRoute(20, 27, riesz_roundtrip=False).apply(x)
    assert rel(a, b) < 1e-12


@pytest.mark.parametrize("N_x,N_t,gamma", CASES[:3])
def test_boundary_rows_are_exactly_zero(N_x, N_t, gamma):
    y = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(rand_x(N_x, N_t)).reshape(2, N_x + 1, N_t)
    assert np.all(y[:, 0, :] == 0) and np.all(y[:, -1, :] == 0)
    yr = DiagFFTPCRefRoute(N_x, N_t, 2.0, gamma).apply(rand_x(N_x, N_t)).reshape(2, N_x + 1, N_t)
    assert np.abs(yr[:, [0, -1], :]).max() == 0


def test_real_input_gives_real_output():
    N_x, N_t = 24, 32
    x = np.random.default_rng(1).standard_normal(2 * (N_x + 1) * N_t)
    y = DiagFFTPCFast(N_x, N_t).apply(x)
    assert np.abs(y.imag).max() < 1e-12 * np.abs(y.real).max()


def test_long_double_route():
    N_x, N_t = 64, 64
    x = rand_x(N_x, N_t)
    f = DiagFFTPCFast(N_x, N_t).apply(x)
    l = DiagFFTPCFast(N_x, N_t, dtype=np.longdouble).apply(x)
    assert l.dtype == np.clongdouble
    assert rel(f, l.astype(complex)) < 1e-11


def test_c_helper_matches_numpy_thomas():
    from oracle import csolve
    N_x, N_t = 96, 40
    x = rand_x(N_x, N_t)
    a = DiagFFTPCFast(N_x, N_t).apply(x)
    b = DiagFFTPCFast(N_x, N_t, solver=csolve.thomas_toeplitz_c).apply(x)
    assert rel(b, a) < 1e-12


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "pc_apply_*.npz"))))
def test_fast_route_matches_golden(path):
    g = np.load(path)
    pc = DiagFFTPCFast(int(g["N_x"]), int(g["N_t"]), float(g["T"]), float(g["gamma"]))
    assert rel(pc.apply(g["x"]), g["y"]) < 1e-11
    assert rel(pc.apply(g["x_real"]), g["y_real"]) < 1e-11


def test_golden_fixtures_exist():
    assert len(glob.glob(os.path.join(GOLDEN, "pc_apply_*.npz"))) >= 4
    assert len(glob.glob(os.path.join(GOLDEN, "gmres_*.npz"))) >= 2


@pytest.mark.parametrize("N_x,N_t,gamma", [(16, 13, 1.0), (40, 64, 1e-2), (33, 20, 1e-4), (257, 128, 1.0)])
def test_threaded_cpu_baseline_equals_the_numpy_route(N_x, N_t, gamma):
    # bench.py's cpu_baseline / reference arm: scipy.fft with workers + the fused pthread stage of oracle/csrc/pc_solve.c
    from oracle.pc_fast import DiagFFTPCFast
    rng = np.random.default_rng(3)
    x = rng.standard_normal(2 * (N_x + 1) * N_t) + 1j * rng.standard_normal(2 * (N_x + 1) * N_t)
    pc = DiagFFTPCFast(N_x, N_t, 2.0, gamma)
    y0, y1 = pc.apply(x), pc.apply_threaded(x)
    assert np.linalg.norm(y1 - y0) <= 1e-11 * np.linalg.norm(y0)      # same recurrence, different complex-division rounding
    assert np.abs(y1.reshape(2, N_x + 1, N_t)[:, [0, -1], :]).max() == 0.0
    # sampled columns of the per-frequency stage (used by the full-size GPU tests), fp64 and 80-bit
    import scipy.fft as sfft
    xh = sfft.ifft(x.reshape(2, N_x + 1, N_t), axis=2)
    ks = np.array([0, 1, N_t // 4, N_t // 2, N_t - 1])
    rp, rm = pc.forward_stage(x)
    zp, zm = pc.solve_stage(rp, rm)
    full = np.stack([zp + zm, (-1j * pc.sigma * pc.z) * (zp - zm)])
    assert np.abs(pc.stage_columns(ks, xh[:, :, ks]) - full[:, :, ks]).max() <= 1e-13 * np.abs(full).max()
    ld = DiagFFTPCFast(N_x, N_t, 2.0, gamma, dtype=np.longdouble).stage_columns(ks, xh[:, :, ks])
    assert np.abs(ld - full[:, :, ks]).max() <= 1e-9 * np.abs(full).max()


def test_lean_gmres_equals_the_reference_restatement():
    from oracle.gmres import gmres, gmres_lean
    from oracle.operator import AllAtOnce
    from oracle.pc_fast import DiagFFTPCFast
    N_x, N_t = 24, 32
    op, pc = AllAtOnce(N_x, N_t), DiagFFTPCFast(N_x, N_t)
    b = np.random.default_rng(0).standard_normal(2 * (N_x + 1) * N_t)
    b = b.reshape(2, N_x + 1, N_t)
    b[:, 0] = b[:, -1] = 0
    b = b.reshape(-1)
    pcr = lambda v: pc.apply(v).real
    x0, i0, h0, r0 = gmres(op.matvec, pcr, b, rtol=1e-8)
    x1, i1, h1, r1 = gmres_lean(op.matvec, pcr, b, rtol=1e-8)
    assert i0 == i1 and r0 == r1 and np.allclose(h0, h1, rtol=1e-10) and np.allclose(x0, x1, rtol=1e-9, atol=1e-12)
    # residual-correction form of the preconditioned operator: same operator, same iteration
    x2, i2, h2, _ = gmres_lean(op.matvec, pcr, b, rtol=1e-8, pc_matvec=lambda v: v + pcr(op.delta(v)))
    assert abs(i2 - i0) <= 1 and np.allclose(x2, x0, rtol=1e-6, atol=1e-9)
    xc = np.random.default_rng(1).standard_normal(b.size).reshape(2, N_x + 1, N_t)
    xc[:, 0] = xc[:, -1] = 0
    assert np.abs(op.delta(xc.reshape(-1)) - (op.matvec(xc.reshape(-1)) - op.pc_matvec(xc.reshape(-1)))).max() < 1e-13
