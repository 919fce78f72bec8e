"""Host mirror of ``Optimal_Control_Wave_Equation`` (Code/Control_Wave_PC.py:13-333) for the
part of it that sits either side of the preconditioner: the right-hand side of the
manufactured problem, the all-at-once operator and the GMRES solve that calls the PC.

Same constructor and method names as upstream (``Build_f``, ``Build_g``,
``Build_Initial_Condition``, ``Build_L``, ``solve(parameters, complex)``); the Firedrake
forms are replaced by the device kernels of libparadiag (``pd_build_rhs``, ``pd_matvec``,
``pd_gmres``).  ``solve(parameters=None)`` is the direct-LU baseline of the ``pc=False`` branch (:186, :573-577).
``write()`` (VTK output, :247-333) is out of scope; ``error_norm`` gives the analytic-solution check it contains
(:299-300, :324-333).
"""
import math
import time

import numpy as np

from .handle import ParaDiagHandle

# solver parameters of the reference run, Control_Wave_PC.py:347-359
default_parameters = {
    'snes_type': 'ksponly',
    'mat_type': 'matfree',
    'ksp_type': 'gmres',
    'ksp_gmres_restart': 300,
    'ksp': {
        'monitor': None,
        'converged_reason': None,
    },
    'ksp_max_it': 1000,
    'pc_type': 'python',
    'pc_python_type': 'optimal_control_paradiag_b200.DiagFFTPC',
}


def _flatten(params, prefix=""):
    out = {}
    for k, v in params.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + k + "_"))
        else:
            out[prefix + k] = v
    return out


class Optimal_Control_Wave_Equation:

    def __init__(self, N_x, T, N_t, gamma, dim=1, device=0, bug138=True):
        if dim != 1:
            raise NotImplementedError("only the 1-D problem is on the accelerated path "
                                      "(upstream dim=2 is a stub, Control_Wave_PC.py:18-19, :296-297)")
        self.N_x = N_x
        self.T = T
        self.N = N_t
        self.gamma = gamma
        self.dt = T / N_t                       # :24
        self.dim = dim
        self.n = N_x + 1
        self.handle = ParaDiagHandle(N_x, N_t, T=T, gamma=gamma, bug138=bug138, device=device)
        self.b = None
        self.ksp_its = None
        self.ksp_reason = None
        self.ksp_history = None

    # The four Build_* of upstream assemble UFL; here they are folded into one kernel that
    # writes b directly (pd_build_rhs), so they only mark which pieces are requested.
    def Build_f(self):
        self._have_f = True                      # :48-61

    def Build_g(self):
        self._have_g = True                      # :65-73

    def Build_Initial_Condition(self):
        self._have_ic = True                     # :76-83

    def Build_L(self):
        """:86-179 -- right-hand side b and the operator A (applied by ``matvec``)."""
        self.b = self.handle.build_rhs()

    def matvec(self, x, y=None):
        return self.handle.matvec(x, y)

    def solve(self, parameters=None, complex=False, rtol=None, verbose=True, real_vectors=False):
        """:182-244.  With the GMRES + python-PC parameters (:347-359) runs the device
        Krylov solve and returns (u_sol, p_sol) as (n, N_t) tensors (node-major, time fastest)."""
        import torch
        if not complex:
            raise NotImplementedError("Should use complex mode (upstream :572, :580)")
        params = _flatten(parameters if parameters else
                          {'ksp_type': 'preonly', 'pc_type': 'lu', 'mat_type': 'aij',
                           'pc_factor_mat_solver_type': 'mumps'})                      # :186
        if params.get('ksp_type') == 'preonly' and params.get('pc_type') == 'lu':
            return self._direct_solve(verbose)                                         # pc=False branch, :573-577
        if params.get('ksp_type') != 'gmres' or params.get('pc_type') != 'python':
            raise NotImplementedError("supported solver configurations: GMRES + python PC (:347-359) and the "
                                      "direct LU baseline preonly + lu (:186)")
        if not str(params.get('pc_python_type', '')).endswith('DiagFFTPC'):
            raise ValueError(f"unknown pc_python_type {params.get('pc_python_type')!r}")
        self.Build_f()
        self.Build_g()
        self.Build_Initial_Condition()
        self.Build_L()
        restart = int(params.get('ksp_gmres_restart', 30))
        max_it = int(params.get('ksp_max_it', 10000))
        # Firedrake's default ksp_rtol is 1e-7 when the options do not set one
        rtol = float(params.get('ksp_rtol', 1e-7)) if rtol is None else rtol
        atol = float(params.get('ksp_atol', 1e-50))
        solver_setted = time.time()
        if real_vectors:
            # the problem is real: float64 Krylov vectors and the half-spectrum preconditioner
            # (pd_gmres_real); the result is promoted so that callers see the same complex layout
            xr, its, hist, reason = self.handle.gmres_real(self.handle.build_rhs_real(), rtol=rtol, atol=atol,
                                                           restart=restart, max_it=max_it)
            x = xr.to(torch.complex128)
        else:
            x, its, hist, reason = self.handle.gmres(self.b, rtol=rtol, atol=atol, restart=restart, max_it=max_it)
        torch.cuda.synchronize(self.handle.device)
        solver_solved = time.time()
        self.ksp_its, self.ksp_reason, self.ksp_history = its, reason, hist
        if verbose:
            if 'ksp_monitor' in params:
                for i, r in enumerate(hist):
                    print(f"  {i:3d} KSP Residual norm {r:.12e}")
            if 'ksp_converged_reason' in params:
                word = "converged" if reason.startswith("CONVERGED") else "did not converge"
                print(f"  Linear solve {word} due to {reason} iterations {its}")
            print("The CPU time for solving the problem", solver_solved - solver_setted)    # :199
        self.U = x
        X = x.view(2, self.n, self.N)
        return X[0], X[1]                                                                    # :200, :244

    DIRECT_MAX_UNKNOWNS = 16384

    def _direct_solve(self, verbose=True):
        """The reference's ``pc=False`` baseline (:186, :573-577: ``ksp_type preonly``, ``pc_type lu``, MUMPS): a direct
        LU solve of the all-at-once system, for end-to-end validation of the GMRES + PC solve at the small sizes the
        upstream accuracy study uses (N_x = N_t = 5 ... 70, plot.py:5-18).  The matrix is assembled on the device
        column by column from the matrix-free operator (``pd_matvec_real``: the problem is real) and factorised by
        cuSOLVER through ``torch.linalg.solve`` -- a library LU standing in for MUMPS; it is a validation baseline,
        not a hot path (dense: limited to DIRECT_MAX_UNKNOWNS unknowns).  Upstream assembles the UNSCALED formulation
        for this branch (:124-133); the scaled system solved here has the same solution after the sqrt(gamma)
        unscaling that ``error_norm`` applies."""
        import torch
        h = self.handle
        sz = h.size
        if sz > self.DIRECT_MAX_UNKNOWNS:
            raise NotImplementedError(f"the dense direct-LU baseline is limited to {self.DIRECT_MAX_UNKNOWNS} unknowns "
                                      f"(got {sz}); use the GMRES + DiagFFTPC parameters of :347-359")
        dev = f"cuda:{h.device}"
        solver_setted = time.time()
        b = h.build_rhs_real()
        A = torch.empty((sz, sz), dtype=torch.float64, device=dev)         # row-major; filled column by column
        e = torch.zeros(sz, dtype=torch.float64, device=dev)
        col = torch.empty(sz, dtype=torch.float64, device=dev)
        for j in range(sz):
            e[j] = 1.0
            h.matvec_real(e, col)
            A[:, j] = col
            e[j] = 0.0
        x = torch.linalg.solve(A, b)
        torch.cuda.synchronize(h.device)
        solver_solved = time.time()
        self.ksp_its, self.ksp_reason, self.ksp_history = 1, "CONVERGED_ITS", []
        if verbose:
            print("The CPU time for solving the problem", solver_solved - solver_setted)    # :199
        self.U = x.to(torch.complex128)
        X = self.U.view(2, self.n, self.N)
        return X[0], X[1]

    def error_norm(self, u_sol):
        """max over time levels of the nodal 2-norm error of u against the analytic state
        sin(pi x) cos(pi t) the data were manufactured from (write(), :299, :324-333), with
        u_sol[:, i] taken at t = (i+1) dt and unscaled by sqrt(gamma) (:289)."""
        u = u_sol.detach().cpu().numpy().real / math.sqrt(self.gamma)
        xs = np.arange(self.n) / self.N_x
        t = (np.arange(self.N) + 1) * self.dt
        ana = np.outer(np.sin(np.pi * xs), np.cos(np.pi * t))
        return float(np.max(np.linalg.norm(u - ana, axis=0)))
