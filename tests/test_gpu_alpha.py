"""The alpha EXTENSION on the GPU (pd_config.alpha != 1) against the oracle's definition of it
(oracle/pc_alpha.py) -- GPU box only.  Parity unpinned by construction: the upstream operator is the
alpha = 1 block circulant and has no alpha; these tests pin the CUDA path to the oracle's three-route
restatement of the generalisation that BASELINE config 5 sweeps."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from optimal_control_paradiag_b200 import DiagFFTPC, ParaDiagError, ParaDiagHandle, _lib, petsc_shim  # noqa: E402
from oracle.gmres import gmres as oracle_gmres  # noqa: E402
from oracle.operator import AllAtOnce  # noqa: E402
from oracle.pc_alpha import DiagFFTPCAlpha, ExplicitAlphaPC  # noqa: E402
from oracle.pc_fast import DiagFFTPCFast  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def rand_x(size, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(size) + 1j * rng.standard_normal(size)


def test_gamma_scaling_kernel():
    N_t, nl, alpha = 96, 11, 1e-3
    x = rand_x(nl * N_t).reshape(nl, N_t)
    g = alpha ** (np.arange(N_t) / N_t)
    with ParaDiagHandle(8, N_t, alpha=alpha) as h:
        xt = torch.tensor(x, device=DEV).reshape(-1)
        yt = torch.empty_like(xt)
        h.stage_gamma(xt, yt, nl, False)
        assert rel(yt.cpu().numpy().reshape(nl, N_t), x * g) < 1e-15
        h.stage_gamma(yt, yt, nl, True)                              # in place, round trip
        assert rel(yt.cpu().numpy().reshape(nl, N_t), x) < 1e-15


# generic N_t, power-of-two N_t (register FFT), N_t divisible by 4, several partition depths in N_x
@pytest.mark.parametrize("N_x,N_t,gamma", [(16, 13, 1.0), (12, 16, 1.0), (20, 32, 1e-2), (80, 81, 1.0), (100, 128, 1.0),
                                           (300, 64, 1e-4), (1024, 256, 1.0)])
@pytest.mark.parametrize("alpha", [0.5, 1e-1, 1e-2, 1e-4, 1e-6])
def test_alpha_apply_matches_oracle(N_x, N_t, gamma, alpha):
    with ParaDiagHandle(N_x, N_t, gamma=gamma, alpha=alpha) as h:
        x = rand_x(h.size)
        ref = DiagFFTPCAlpha(N_x, N_t, 2.0, gamma, alpha).apply(x)
        y = h.pc_apply(torch.tensor(x, device=DEV)).cpu().numpy()
        # P_alpha carries 1/alpha entries: its conditioning grows as alpha -> 0 (two fp64 routes differ by
        # cond * eps); 1e-10 is asserted down to alpha = 1e-2
        assert rel(y, ref) < (1e-10 if alpha >= 1e-2 else 1e-7)
        assert np.abs(y.reshape(2, N_x + 1, N_t)[:, [0, -1], :]).max() == 0.0
        assert np.array_equal(h.pc_apply_host(x), y)
        xt = torch.tensor(x, device=DEV)
        h.pc_apply(xt, xt)                                           # in place
        assert np.array_equal(xt.cpu().numpy(), y)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "alpha_*.npz"))))
def test_alpha_apply_matches_explicit_matrix_golden(path):
    g = np.load(path)
    with ParaDiagHandle(int(g["N_x"]), int(g["N_t"]), T=float(g["T"]), gamma=float(g["gamma"]),
                        alpha=float(g["alpha"])) as h:
        y = h.pc_apply(torch.tensor(g["x"], device=DEV)).cpu().numpy()
        assert rel(y, g["y"]) < 1e-10


def test_alpha_to_one_is_continuous():
    N_x, N_t = 64, 81
    x = rand_x(2 * (N_x + 1) * N_t)
    with ParaDiagHandle(N_x, N_t) as h1, ParaDiagHandle(N_x, N_t, alpha=1 - 1e-9) as ha:
        y1 = h1.pc_apply(torch.tensor(x, device=DEV)).cpu().numpy()
        ya = ha.pc_apply(torch.tensor(x, device=DEV)).cpu().numpy()
        assert rel(ya, y1) < 1e-6


@pytest.mark.parametrize("alpha,expected", [(0.5, 12), (1e-1, 18), (1e-2, 25)])
def test_alpha_gmres_iteration_counts(alpha, expected):
    # SURVEY H1 / tests/test_oracle_alpha.py: N_x = 20, N_t = 32, gamma = 1, rtol 1e-7
    N_x, N_t = 20, 32
    op = AllAtOnce(N_x, N_t)
    pc = DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha)
    _, its_o, hist_o, _ = oracle_gmres(op.matvec, pc.apply, op.rhs() + 0j, rtol=1e-7)
    assert its_o == expected
    with ParaDiagHandle(N_x, N_t, alpha=alpha) as h:
        x, its, hist, reason = h.gmres(h.build_rhs(), rtol=1e-7)
        assert reason == "CONVERGED_RTOL" and abs(its - its_o) <= 1
        assert np.allclose(hist[:5], hist_o[:5], rtol=1e-6)


def test_alpha_through_the_pc_class_and_options():
    N_x, N_t, alpha = 40, 48, 1e-2
    x = rand_x(2 * (N_x + 1) * N_t)
    ref = DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha).apply(x)
    DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=1.0)
    try:
        pc = petsc_shim.PC()
        pc.setOptionsPrefix("fieldsplit_")
        pc.options["fieldsplit_diagfft_alpha"] = str(alpha)
        pc.setPythonContext(DiagFFTPC())
        pc.setUp()
        xv, yv = petsc_shim.Vec(x), petsc_shim.Vec.zeros(x.size)
        pc.apply(xv, yv)
        assert rel(yv.getArray(), ref) < 1e-10
        pc.destroy()
    finally:
        DiagFFTPC._defaults = {}


# ---- alpha on the real-input (half-spectrum) path: Gamma is real and lambda(N_t - k) = conj lambda(k) still holds
@pytest.mark.parametrize("N_x,N_t,gamma", [(100, 128, 1.0), (300, 256, 1e-2), (64, 1024, 1.0), (1024, 512, 1.0),
                                           (37, 4096, 1.0)])
@pytest.mark.parametrize("alpha", [0.5, 1e-2, 1e-4])
def test_alpha_real_input_path_matches_oracle(N_x, N_t, gamma, alpha):
    with ParaDiagHandle(N_x, N_t, gamma=gamma, alpha=alpha) as h:
        assert h.real_path_supported
        x = np.random.default_rng(3).standard_normal(h.size)
        ref = DiagFFTPCAlpha(N_x, N_t, 2.0, gamma, alpha).apply(x + 0j)
        assert np.abs(ref.imag).max() <= 1e-9 * np.abs(ref.real).max()      # P_alpha^-1 of a real vector is real
        xt = torch.tensor(x, device=DEV)
        y = h.pc_apply_real(xt).cpu().numpy()
        assert rel(y, ref.real) < (1e-10 if alpha >= 1e-2 else 1e-7)
        assert np.abs(y.reshape(2, N_x + 1, N_t)[:, [0, -1], :]).max() == 0.0
        yc = h.pc_apply(torch.tensor(x + 0j, device=DEV)).cpu().numpy()      # the complex path of the same handle
        assert rel(y, yc.real) < (1e-11 if alpha >= 1e-2 else 1e-8)


@pytest.mark.parametrize("N_x,N_t", [(20, 16384), (700, 8192)])
def test_alpha_real_input_path_large_nt(N_x, N_t):
    # N_t = 16384 takes the per-line packed real transform (Gamma on the interleaved samples), 8192 the pair kernel
    alpha = 1e-2
    with ParaDiagHandle(N_x, N_t, alpha=alpha) as h:
        x = np.random.default_rng(4).standard_normal(h.size)
        y = h.pc_apply_real(torch.tensor(x, device=DEV)).cpu().numpy()
        yc = h.pc_apply(torch.tensor(x + 0j, device=DEV)).cpu().numpy()
        assert rel(y, yc.real) < 1e-11
        if N_x == 20:
            ref = DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha).apply(x + 0j)
            assert rel(y, ref.real) < 1e-10


def test_alpha_real_gmres_matches_complex_gmres():
    N_x, N_t, alpha = 64, 128, 1e-1
    with ParaDiagHandle(N_x, N_t, alpha=alpha) as h:
        xc, its_c, hist_c, reason_c = h.gmres(h.build_rhs(), rtol=1e-7)
        xr, its_r, hist_r, reason_r = h.gmres_real(h.build_rhs_real(), rtol=1e-7)
        assert reason_c == reason_r == "CONVERGED_RTOL" and abs(its_c - its_r) <= 1
        assert rel(xr.cpu().numpy(), xc.cpu().numpy().real) < 1e-6


# ---- alpha in slab mode (x-slab sharding, peer-store exchange): against the single-GPU alpha apply and the oracle
@pytest.mark.parametrize("N_x,N_t,G", [(80, 81, 2), (255, 128, 3), (1024, 256, 4), (1024, 1024, 8), (40, 16384, 2),
                                        (4096, 64, 8)])
@pytest.mark.parametrize("alpha", [0.5, 1e-2])
def test_alpha_slab_mode_equals_single_gpu(N_x, N_t, G, alpha):
    from optimal_control_paradiag_b200.dist import LocalSlabGroup
    with ParaDiagHandle(N_x, N_t, alpha=alpha) as h, LocalSlabGroup(N_x, N_t, G, alpha=alpha) as grp:
        for rep in range(3):
            x = torch.tensor(rand_x(h.size, seed=rep), device=DEV)
            ref = h.pc_apply(x)
            y = grp.apply(x)
            err = float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref))
            assert err < 1e-10, (rep, err)
            assert float(y.view(2, N_x + 1, N_t)[:, [0, -1], :].abs().max()) == 0.0
        if h.real_path_supported:
            xr = torch.tensor(np.random.default_rng(7).standard_normal(h.size), device=DEV)
            ref = h.pc_apply_real(xr)
            y = grp.apply(xr, real=True)
            assert float(torch.linalg.norm(y - ref) / torch.linalg.norm(ref)) < 1e-10
        assert all(not to for to, _ in grp.status())
    if N_x * N_t <= 300000:
        x = rand_x(2 * (N_x + 1) * N_t, seed=5)
        with LocalSlabGroup(N_x, N_t, G, alpha=alpha) as grp:
            y = grp.apply(torch.tensor(x, device=DEV)).cpu().numpy()
        # (two fp64 routes differ by cond * eps, growing with N_x: 4e-10 at N_x = 4096 on one GPU as well)
        assert rel(y, DiagFFTPCAlpha(N_x, N_t, 2.0, 1.0, alpha).apply(x)) < (1e-10 if N_x <= 1024 else 2e-9)


def test_alpha_unsupported_paths_fail_loudly():
    with ParaDiagHandle(16, 128, alpha=0.1) as h:
        with pytest.raises(ParaDiagError) as ei:
            h.pc_matvec(torch.zeros(h.size, dtype=torch.complex128, device=DEV))
        assert ei.value.status == _lib.PD_ERR_UNSUPPORTED
    with pytest.raises(ParaDiagError) as ei:                     # frequency-sharded stage handles (all-to-all mode)
        ParaDiagHandle(16, 16, alpha=0.1, k_begin=0, k_count=8)
    assert ei.value.status == _lib.PD_ERR_UNSUPPORTED
    with pytest.raises(ParaDiagError) as ei:
        ParaDiagHandle(16, 16, alpha=2.0)
    assert ei.value.status == _lib.PD_ERR_INVALID
