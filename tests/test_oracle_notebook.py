"""Known-answer identities of the upstream notebook Code/mat_test.ipynb (the only checks the
reference holds for this path) re-run against the oracle's conventions."""
import numpy as np
import pytest
import scipy.linalg as sla
from scipy.fft import fft, ifft

from oracle import eigs


@pytest.mark.parametrize("N_t", [5, 64, 81])
def test_cells_5_to_9_fft_convention_and_circulants(N_t):
    l1, l2 = eigs.lambdas(N_t)
    It = np.eye(N_t)
    A = ifft(np.diag(l1).conj() @ fft(It, axis=0), axis=0)          # cell 5
    B = fft(np.diag(l2) @ ifft(It, axis=0), axis=0)                 # cell 5
    Cm = ifft(np.diag(l2).conj() @ fft(It, axis=0), axis=0)         # cell 6
    assert np.linalg.norm(B - Cm) < 1e-13                           # cell 7: 1.1447e-15 at N_t=5
    first_col = np.zeros(N_t)
    first_col[:3] = [1, -2, 1]
    assert np.linalg.norm(A - sla.circulant(first_col)) < 1e-13     # cell 9: 8.4159e-16 at N_t=5


@pytest.mark.parametrize("N_t", [5, 64, 81])
def test_cells_1_2_11_12_diagonalisation(N_t):
    T, gamma = 2, 1
    tau = T / N_t
    l1, l2 = eigs.lambdas(N_t)
    with np.errstate(all="ignore"):
        S1 = np.sqrt(-np.conj(l2) / l2)
        S2 = -np.sqrt(-l2 / np.conj(l2))
        m1 = np.real(l1 / l2)
        m2 = -tau ** 2 / np.conj(l2) / np.sqrt(gamma)
        m3 = tau ** 2 / l2 / np.sqrt(gamma)
    ok = np.abs(l2) > 1e-12          # N_t % 4 == 0 has lambda_2 = 0 at two frequencies
    Sigma_1 = m1 + m2 * S1
    Sigma_2 = m1 + m3 * S2
    for k in np.nonzero(ok)[0]:
        S = np.array([[1, S2[k]], [S1[k], 1]])
        Lam = np.array([[m1[k], m2[k]], [m3[k], m1[k]]])
        Sig = np.diag([Sigma_1[k], Sigma_2[k]])
        assert np.allclose(S @ S.conj().T, 2 * np.eye(2), atol=1e-12)        # cell 2
        assert np.linalg.norm(Lam @ S - S @ Sig) < 1e-10 * max(1, np.abs(Lam).max())  # cell 12


def test_closed_forms_match_numpy_eig():
    for N_t, gamma in [(81, 1.0), (64, 1e-2), (100, 1e-4)]:
        dt = 2.0 / N_t
        Sig, S, SI = eigs.eig_numpy(N_t, dt, gamma)
        l1, l2 = eigs.lambdas(N_t)
        cf = eigs.closed_form(N_t, dt, gamma)
        # lambda_2 = 2 cos(theta) z, lambda_1 = -4 sin^2(theta/2) z
        assert np.allclose(2 * cf["kappa"] / dt ** 2 * cf["z"], l2, atol=1e-14)
        assert np.allclose(cf["s_re"] * cf["z"], l1, atol=1e-14)
        ok = np.abs(l2) > 1e-9
        sig_pm = np.real(l1[ok] / l2[ok])[:, None] + 1j * cf["c"] / np.abs(l2[ok])[:, None] * np.array([1, -1])
        got = np.sort_complex(Sig[ok].round(12))
        want = np.sort_complex(sig_pm.round(12))
        for g_, w_ in zip(Sig[ok], sig_pm):
            assert min(abs(g_[0] - w_[0]) + abs(g_[1] - w_[1]), abs(g_[0] - w_[1]) + abs(g_[1] - w_[0])) \
                < 1e-9 * max(1.0, abs(w_[0]))
        # eigenvector matrices are unitary up to column scaling: cond(S) = 1
        for k in np.nonzero(ok)[0]:
            assert np.linalg.cond(S[k]) < 1 + 1e-8
