"""All-at-once operator and manufactured right-hand side of the reference problem.

Oracle code (test infrastructure).  Restates, for the ``pc=True`` branches,
``Optimal_Control_Wave_Equation.Build_L`` (Control_Wave_PC.py:86-179) as the
linear operator ``A`` the matrix-free Jacobian applies, and ``Build_f`` /
``Build_g`` / ``Build_Initial_Condition`` (:48-83) as the right-hand side ``b``
(``snes_type ksponly`` with zero initial guess solves ``A U = -F(0) = b``).

Layout: ``x.reshape(2, n, N_t)`` (field, node, time-fastest), as in the PC.

With M, K the interior P1 mass/stiffness blocks (Dirichlet columns dropped,
:44-45), c = dt^2/sqrt(gamma):

  state row i   : M(u_i - 2u_{i-1} + u_{i-2}) + q_i dt^2/2 K(u_i + u_{i-2}) - d_i c M p_i
                  u_{-1} = u_{-2} = 0 (known initial data moves to b, :93-95, :118)
                  d_0 = 1/2 (:117), q_{N-1} = sqrt(gamma) (the :138 quirk, ``bug138``)
  adjoint row i : e_i c M u_i + M(p_i - 2p_{i+1} + p_{i+2}) + dt^2/2 K(p_i + p_{i+2})
                  p_N = p_{N+1} = 0 (:102-107), e_{N-1} = 1/2 (:143), and the last
                  row carries no second difference (:141-142)
  boundary rows : identity.
"""
import numpy as np

from . import fem1d


class AllAtOnce:
    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, bug138=True):
        self.N_x, self.N_t, self.n = N_x, N_t, N_x + 1
        self.T, self.gamma, self.bug138 = T, gamma, bug138
        self.dt = T / N_t
        self.c = self.dt ** 2 / np.sqrt(gamma)
        self.Mf = fem1d.mass_full(N_x)
        self.Kf = fem1d.stiff_full(N_x)

    # interior rows of M v / K v with boundary columns dropped
    def _M(self, v):
        w = v.copy()
        w[0] = w[-1] = 0
        out = self.Mf @ w
        out[0] = out[-1] = 0
        return out

    def _K(self, v):
        w = v.copy()
        w[0] = w[-1] = 0
        out = self.Kf @ w
        out[0] = out[-1] = 0
        return out

    def matvec(self, x):
        n, N = self.n, self.N_t
        x = np.asarray(x).reshape(2, n, N)
        u, p = x[0], x[1]
        dt2h = self.dt ** 2 / 2
        Mu, Ku, Mp, Kp = self._M(u), self._K(u), self._M(p), self._K(p)
        y = np.zeros_like(x)
        d = np.ones(N)
        d[0] = 0.5                                            # :117
        e = np.ones(N)
        e[-1] = 0.5                                           # :143
        q = np.ones(N)
        if self.bug138:
            q[-1] = np.sqrt(self.gamma)                       # :138
        yu = Mu.copy()
        yu[:, 1:] -= 2 * Mu[:, :-1]
        yu[:, 2:] += Mu[:, :-2]
        ku = Ku.copy()
        ku[:, 2:] += Ku[:, :-2]
        yu += dt2h * ku * q
        yu -= self.c * Mp * d
        yp = Mp.copy()
        yp[:, :-1] -= 2 * Mp[:, 1:]
        yp[:, :-2] += Mp[:, 2:]
        kp = Kp.copy()
        kp[:, :-2] += Kp[:, 2:]
        yp += dt2h * kp
        yp += self.c * Mu * e
        y[0], y[1] = yu, yp
        y[:, 0, :] = x[:, 0, :]
        y[:, -1, :] = x[:, -1, :]
        return y.reshape(-1)

    def pc_matvec(self, x):
        """y = P x for the block-circulant matrix ``DiagFFTPC`` inverts: the operator above with the time stencils
        made periodic (C1, C2 of mat_test.ipynb cells 8-9) and the factors d_0, e_{N-1}, q_{N-1} replaced by 1."""
        n, N = self.n, self.N_t
        x = np.asarray(x).reshape(2, n, N)
        u, p = x[0], x[1]
        dt2h = self.dt ** 2 / 2
        Mu, Ku, Mp, Kp = self._M(u), self._K(u), self._M(p), self._K(p)
        y = np.zeros_like(x)
        y[0] = Mu - 2 * np.roll(Mu, 1, axis=1) + np.roll(Mu, 2, axis=1) + dt2h * (Ku + np.roll(Ku, 2, axis=1)) - self.c * Mp
        y[1] = Mp - 2 * np.roll(Mp, -1, axis=1) + np.roll(Mp, -2, axis=1) + dt2h * (Kp + np.roll(Kp, -2, axis=1)) + self.c * Mu
        y[:, 0, :] = x[:, 0, :]
        y[:, -1, :] = x[:, -1, :]
        return y.reshape(-1)

    def delta(self, x):
        """(A - P) x written out term by term -- the residual-correction operator (pd_delta): only the wrap-around
        time levels and the special factors (:117, :143, :138) differ between A and P, so the result is non-zero on
        at most three time levels per field and no second difference is ever formed."""
        n, N = self.n, self.N_t
        x = np.asarray(x).reshape(2, n, N)
        u, p = x[0], x[1]
        dt2h = self.dt ** 2 / 2
        d = np.zeros_like(x)
        Mc = lambda v: self._M(v.reshape(n, 1))[:, 0]
        Kc = lambda v: self._K(v.reshape(n, 1))[:, 0]
        for i in range(min(2, N)):
            if i - 1 < 0:
                d[0][:, i] += 2 * Mc(u[:, i - 1 + N])
            if i - 2 < 0:
                d[0][:, i] -= Mc(u[:, i - 2 + N]) + dt2h * Kc(u[:, i - 2 + N])
        d[0][:, 0] += 0.5 * self.c * Mc(p[:, 0])
        if self.bug138 and self.gamma != 1.0:
            k2 = Kc(u[:, N - 1]) + (Kc(u[:, N - 3]) if N - 3 >= 0 else 0)
            d[0][:, N - 1] += (np.sqrt(self.gamma) - 1.0) * dt2h * k2
        for i in range(max(N - 2, 0), N):
            if i + 1 >= N:
                d[1][:, i] += 2 * Mc(p[:, i + 1 - N])
            if i + 2 >= N:
                d[1][:, i] -= Mc(p[:, i + 2 - N]) + dt2h * Kc(p[:, i + 2 - N])
        d[1][:, N - 1] -= 0.5 * self.c * Mc(u[:, N - 1])
        d[:, 0, :] = 0
        d[:, -1, :] = 0
        return d.reshape(-1)

    def rhs(self):
        """b of ``A U = b`` for the manufactured data (:48-83, pc=True scaling)."""
        n, N, dt, T, gamma = self.n, self.N_t, self.dt, self.T, self.gamma
        xs = np.arange(n) / self.N_x
        sx = np.sin(np.pi * xs)
        sg = np.sqrt(gamma)
        ti = np.arange(N) * dt
        f = (-1.0 / gamma) * np.outer(sx, (np.exp(ti) - np.exp(T)) ** 2) * sg   # :55-57
        tg = np.arange(1, N + 1) * dt
        g = np.outer(sx, 2 * (2 * np.exp(2 * tg) - np.exp(T + tg))
                     + np.pi ** 2 * (np.exp(tg) - np.exp(T)) ** 2
                     + np.cos(np.pi * tg))                                        # :70-72
        u0 = sg * sx                                                               # :79
        u1 = np.zeros(n)                                                           # :80
        M = lambda v: self._interior_rows(self.Mf @ v)
        K = lambda v: self._interior_rows(self.Kf @ v)
        b = np.zeros((2, n, N))
        b[0] = dt ** 2 * M(f)                                                      # :139, :159
        b[0][:, 0] = dt ** 2 * M(0.5 * f[:, 0] + u1 / dt + u0 / dt ** 2)           # :118
        if N > 1:
            b[0][:, 1] += -M(u0) - dt ** 2 / 2 * K(u0)                             # :93-95, :157-158
        b[1] = dt ** 2 * M(g)                                                      # :123, :164
        b[1][:, N - 1] *= 0.5                                                      # :144
        return b.reshape(-1)

    @staticmethod
    def _interior_rows(v):
        v = np.array(v, copy=True)
        v[0] = v[-1] = 0
        return v

    def dense(self):
        """Dense A (tiny sizes only), column by column."""
        sz = 2 * self.n * self.N_t
        A = np.zeros((sz, sz))
        I = np.eye(sz)
        for j in range(sz):
            A[:, j] = self.matvec(I[:, j])
        return A

    def direct_solve(self, b=None):
        """The reference's ``pc=False`` baseline (:186, :573-577: ``ksp_type preonly``, ``pc_type lu``, MUMPS):
        a direct LU solve of the all-at-once system, here SuperLU on the explicit matrix (small sizes)."""
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        b = self.rhs() if b is None else np.asarray(b)
        lu = spla.splu(sp.csc_matrix(self.dense()))
        if np.iscomplexobj(b):
            return lu.solve(b.real) + 1j * lu.solve(b.imag)
        return lu.solve(b)

    def analytic(self):
        """Nodal values of the analytic state/adjoint the data were manufactured
        from (write(), :299-300), at the times the unknowns live on:
        u_i ~ u((i+1) dt) * sqrt(gamma),  p_i ~ p(i dt)."""
        n, N, dt, T = self.n, self.N_t, self.dt, self.T
        xs = np.arange(n) / self.N_x
        sx = np.sin(np.pi * xs)
        tu = np.arange(1, N + 1) * dt
        tp = np.arange(N) * dt
        u = np.sqrt(self.gamma) * np.outer(sx, np.cos(np.pi * tu))
        p = np.outer(sx, (np.exp(tp) - np.exp(T)) ** 2)
        return u, p
