"""Left-preconditioned restarted GMRES with PETSc KSPGMRES semantics.

Oracle code (test infrastructure).  Restates the Krylov solve that the options
of Control_Wave_PC.py:347-359 select (``ksp_type gmres``, ``ksp_gmres_restart
300``, ``ksp_max_it 1000``, ``pc_type python``; everything else PETSc default):

* left preconditioning, zero initial guess (``snes_type ksponly``),
* classical Gram-Schmidt, no refinement (PETSc default
  ``KSPGMRESClassicalGramSchmidtOrthogonalization`` with refine never),
* convergence tested on the *preconditioned* residual norm
  ``|| P^-1 (b - A x_k) ||_2 <= max(rtol * || P^-1 b ||_2, atol)``, estimated
  from the Givens-rotated least-squares right-hand side,
* the iteration count is the number of Krylov steps (matvec + PC apply pairs),
  which is what ``-ksp_monitor`` / ``-ksp_converged_reason`` print (:352-355).

The inner products follow PETSc's ``VecMDot``: ``h_i = v_i^H w``.
"""
import numpy as np


def gmres(matvec, pc_apply, b, rtol=1e-5, atol=1e-50, restart=300, max_it=1000,
          monitor=None, pc_matvec=None):
    """``pc_matvec`` (optional): v -> P^-1 A v evaluated some other way (the residual-correction form
    v + P^-1 (A - P) v of pd_set_option "gmres_residual_correction"); the rest of the iteration is unchanged."""
    b = np.asarray(b)
    its = 0
    hist = []
    r = pc_apply(b)
    x = np.zeros_like(r)
    beta0 = np.linalg.norm(r)
    target = max(rtol * beta0, atol)
    hist.append(beta0)
    if monitor:
        monitor(0, beta0)
    if beta0 <= target or beta0 == 0.0:
        return x, 0, hist, "CONVERGED_ATOL" if beta0 <= atol else "CONVERGED_RTOL"
    reason = "DIVERGED_ITS"
    while its < max_it:
        if its > 0:
            r = pc_apply(b - matvec(x))
        beta = np.linalg.norm(r)
        m = restart
        V = [r / beta]
        H = np.zeros((m + 1, m), dtype=r.dtype)
        cs = np.zeros(m, dtype=r.dtype)
        sn = np.zeros(m, dtype=r.dtype)
        g = np.zeros(m + 1, dtype=r.dtype)
        g[0] = beta
        j_done = 0
        converged = False
        for j in range(m):
            w = pc_matvec(V[j]) if pc_matvec is not None else pc_apply(matvec(V[j]))
            Vm = np.array(V)
            h = Vm.conj() @ w                      # classical Gram-Schmidt: all dots first
            w = w - h @ Vm
            H[: j + 1, j] = h
            hn = np.linalg.norm(w)
            H[j + 1, j] = hn
            # apply previous rotations
            for i in range(j):
                t = np.conj(cs[i]) * H[i, j] + np.conj(sn[i]) * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            a_, b_ = H[j, j], H[j + 1, j]
            den = np.sqrt(abs(a_) ** 2 + abs(b_) ** 2)
            if den == 0:
                cs[j], sn[j] = 1.0, 0.0
            else:
                cs[j], sn[j] = a_ / den, b_ / den
            H[j, j] = np.conj(cs[j]) * a_ + np.conj(sn[j]) * b_
            H[j + 1, j] = 0
            g[j + 1] = -sn[j] * g[j]
            g[j] = np.conj(cs[j]) * g[j]
            its += 1
            j_done = j + 1
            rn = abs(g[j + 1])
            hist.append(rn)
            if monitor:
                monitor(its, rn)
            if rn <= target:
                converged = True
                reason = "CONVERGED_RTOL" if rn > atol else "CONVERGED_ATOL"
                break
            if its >= max_it:
                break
            if hn == 0:
                break
            V.append(w / hn)
        # solve the small triangular system and update x
        yk = np.linalg.solve(np.triu(H[:j_done, :j_done]), g[:j_done]) if j_done else np.zeros(0)
        for i in range(j_done):
            x = x + yk[i] * V[i]
        if converged or its >= max_it:
            break
    return x, its, hist, reason


def gmres_lean(matvec, pc_apply, b, rtol=1e-5, atol=1e-50, restart=300, max_it=1000, monitor=None, pc_matvec=None):
    """Same iteration as ``gmres`` (identical arithmetic order per step: classical Gram-Schmidt with all inner
    products taken against the unmodified w, Givens-rotated residual estimate) with the Krylov basis held in
    ONE growing 2-D array instead of a list that is re-stacked every step -- for the BASELINE-size golden
    runs (tests/golden/make_gmres_counts.py), where a basis vector is 0.27-1.1 GB."""
    b = np.asarray(b)
    its, hist = 0, []
    r = pc_apply(b)
    x = np.zeros_like(r)
    beta0 = float(np.linalg.norm(r))
    target = max(rtol * beta0, atol)
    hist.append(beta0)
    if monitor:
        monitor(0, beta0)
    if beta0 <= target or beta0 == 0.0:
        return x, 0, hist, "CONVERGED_ATOL" if beta0 <= atol else "CONVERGED_RTOL"
    reason = "DIVERGED_ITS"
    V = np.empty((8, r.size), dtype=r.dtype)
    while its < max_it:
        if its > 0:
            r = pc_apply(b - matvec(x))
        beta = float(np.linalg.norm(r))
        m = restart
        V[0] = r / beta
        H = np.zeros((m + 1, m), dtype=r.dtype)
        cs = np.zeros(m, dtype=r.dtype)
        sn = np.zeros(m, dtype=r.dtype)
        g = np.zeros(m + 1, dtype=r.dtype)
        g[0] = beta
        j_done, converged = 0, False
        for j in range(m):
            w = pc_matvec(V[j]) if pc_matvec is not None else pc_apply(matvec(V[j]))
            h = V[: j + 1].conj() @ w
            w = w - h @ V[: j + 1]
            H[: j + 1, j] = h
            hn = float(np.linalg.norm(w))
            H[j + 1, j] = hn
            for i in range(j):
                t = np.conj(cs[i]) * H[i, j] + np.conj(sn[i]) * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            a_, b_ = H[j, j], H[j + 1, j]
            den = np.sqrt(abs(a_) ** 2 + abs(b_) ** 2)
            if den == 0:
                cs[j], sn[j] = 1.0, 0.0
            else:
                cs[j], sn[j] = a_ / den, b_ / den
            H[j, j] = np.conj(cs[j]) * a_ + np.conj(sn[j]) * b_
            H[j + 1, j] = 0
            g[j + 1] = -sn[j] * g[j]
            g[j] = np.conj(cs[j]) * g[j]
            its += 1
            j_done = j + 1
            rn = abs(g[j + 1])
            hist.append(float(rn))
            if monitor:
                monitor(its, rn)
            if rn <= target:
                converged = True
                reason = "CONVERGED_RTOL" if rn > atol else "CONVERGED_ATOL"
                break
            if its >= max_it or hn == 0:
                break
            if j + 2 > V.shape[0]:
                V = np.concatenate([V, np.empty((min(V.shape[0], 16), r.size), dtype=r.dtype)])
            V[j + 1] = w / hn
        yk = np.linalg.solve(np.triu(H[:j_done, :j_done]), g[:j_done]) if j_done else np.zeros(0)
        for i in range(j_done):
            x = x + yk[i] * V[i]
        if converged or its >= max_it:
            break
    return x, its, hist, reason
