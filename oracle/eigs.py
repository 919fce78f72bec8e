"""Time-circulant eigenvalues and the per-frequency 2x2 diagonalisation.

Oracle code (test infrastructure).  Follows Control_Wave_PC.py:

* ``lambdas``      :387-388  lambda_1(k) = 1 - 2 z + z^2, lambda_2(k) = 1 + z^2,
                   z = exp(2 pi i k / N_t).
* ``Lambda_block`` :418-419  the 2x2 matrix whose eigen-decomposition the
                   upstream loop :415-436 takes with numpy.
* ``eig_numpy``    :421-425  S, Sigma, S^-1 exactly as upstream computes them
                   (``np.linalg.eig`` + ``np.linalg.inv``).
* ``closed_form``  the analytic expressions of the abandoned draft
                   pre_cond.py:32-38 / mat_test.ipynb cell 1, in the
                   division-free shape the CUDA kernels regenerate:
                   lambda_2 = 2 cos(t) e^{it}, lambda_1 = -4 sin^2(t/2) e^{it},
                   e^{i phi} = sign(cos t) e^{it}, Sigma_pm = Re(l1/l2) +- i c/|l2|.
"""
import numpy as np


def lambdas(N_t):
    k = np.arange(N_t)
    l1 = 1 - 2 * np.exp(2j * np.pi / N_t * k) + np.exp(4j * np.pi / N_t * k)
    l2 = 1 + np.exp(4j * np.pi / N_t * k)
    return l1, l2


def Lambda_block(l1, l2, dt, gamma):
    c = dt ** 2 / np.sqrt(gamma)
    return np.array([[l1 / l2, -c / np.conj(l2)],
                     [c / l2, np.conj(l1) / np.conj(l2)]])


def eig_numpy(N_t, dt, gamma):
    """Arrays (Sigma_1, Sigma_2, S[4], Sinv[4]) as the loop :415-436 builds them."""
    l1, l2 = lambdas(N_t)
    S = np.zeros((N_t, 2, 2), dtype=complex)
    SI = np.zeros((N_t, 2, 2), dtype=complex)
    Sig = np.zeros((N_t, 2), dtype=complex)
    with np.errstate(all="ignore"):
        for i in range(N_t):
            Lam = Lambda_block(l1[i], l2[i], dt, gamma)
            e, v = np.linalg.eig(Lam)
            Sig[i] = e
            S[i] = v
            SI[i] = np.linalg.inv(v)
    return Sig, S, SI


def angles(N_t, dtype=np.float64):
    """cos/sin of theta_k = 2 pi k / N_t, exact at multiples of pi/2."""
    k = np.arange(N_t)
    # reduce 2k/N_t (half-turns) so that exact zeros of cos/sin are exact
    num = (4 * k) % (4 * N_t)  # theta = num * pi / (2 N_t)
    th = num.astype(dtype) * dtype(np.pi) / dtype(2 * N_t)
    c, s = np.cos(th), np.sin(th)
    quarter = (num % N_t) == 0
    q = (num // N_t) % 4
    c = np.where(quarter, np.array([1, 0, -1, 0], dtype=dtype)[q], c)
    s = np.where(quarter, np.array([0, 1, 0, -1], dtype=dtype)[q], s)
    return c, s


def closed_form(N_t, dt, gamma, dtype=np.float64):
    """Division-free per-frequency quantities.

    Returns dict with
      z      = e^{i theta}
      sigma  = sign(cos theta) (+1 where cos theta == 0)
      s_re   = -4 sin^2(theta/2)            (real part of lambda_1 e^{-i theta})
      kappa  = dt^2 cos(theta)              (= dt^2/2 * lambda_2 e^{-i theta})
      c      = dt^2 / sqrt(gamma)
    so that T_pm(k) = e^{i theta} [ (s_re +- i c sigma) M + kappa K ].
    """
    cth, sth = angles(N_t, dtype)
    ctype = np.result_type(dtype, np.complex64) if dtype != np.longdouble else np.clongdouble
    z = (cth + 1j * sth).astype(ctype)
    sigma = np.where(cth >= 0, dtype(1), dtype(-1))
    half = np.sin(np.arange(N_t).astype(dtype) * dtype(np.pi) / dtype(N_t))
    s_re = -dtype(4) * half * half  # -4 sin^2(theta/2), no cancellation at small k
    kappa = dtype(dt) ** 2 * cth
    c = dtype(dt) ** 2 / np.sqrt(dtype(gamma))
    return dict(z=z, sigma=sigma, s_re=s_re, kappa=kappa, c=c)
