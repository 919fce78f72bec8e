"""Regenerates tests/golden/gmres_counts.json: GMRES iteration counts and residual histories of the CPU
oracle at BASELINE sizes (north star: "GMRES iteration counts match within +-1").

The upstream code cannot run here (no Firedrake / PETSc / MUMPS) and holds no recorded iteration counts, so
these are counts of the oracle's restatement (oracle/gmres.py KSPGMRES semantics, oracle/operator.py Build_L,
oracle/pc_fast.py apply with the threaded C Thomas stage) -- "parity unpinned" like every oracle-derived
fixture.  The problem is real (b, A, P real), so the oracle runs on float64 vectors: the imaginary part the
complex solve carries is rounding noise (tests/test_oracle_operator_gmres.py checks real == complex at small
sizes).  One rtol = 1e-9 run per case yields the whole history; the count for any looser rtol is the first
index whose residual is below rtol * hist[0] (no restart happens within 300 steps).

Each case is also run in the residual-correction form of the preconditioned operator, v + P^-1 (A - P) v
(``correction``): no cancelling second difference is formed, the count is the exact-arithmetic one and does not move
under the perturbation.

Also records, per case, how the count reacts to a relative perturbation eps of the PC output (a stand-in for a
DIFFERENT but equally valid fp64 implementation of the same preconditioner): where the count moves, it is set
by rounding and a +-1 comparison between implementations is not meaningful.

Run from the repo root (cfg5 needs ~15 GB of RAM and a few minutes):
    python tests/golden/make_gmres_counts.py [case ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.gmres import gmres_lean  # noqa: E402
from oracle.operator import AllAtOnce  # noqa: E402
from oracle.pc_fast import DiagFFTPCFast  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "gmres_counts.json")
CASES = {
    "cfg1": (80, 81, 1.0), "cfg2": (1024, 1024, 1.0), "mid": (4096, 512, 1.0), "cfg5": (4096, 4096, 1.0),
    "cfg5_g1e-2": (4096, 4096, 1e-2), "cfg5_g1e-4": (4096, 4096, 1e-4), "cfg3": (16384, 4096, 1.0),
}


def count(hist, rtol):
    for i, r in enumerate(hist):
        if r <= rtol * hist[0]:
            return i
    return None


def run(N_x, N_t, gamma, eps=0.0, rtol=1e-9, max_it=60, seed=7, correction=False):
    op = AllAtOnce(N_x, N_t, 2.0, gamma)
    pc = DiagFFTPCFast(N_x, N_t, 2.0, gamma)
    rng = np.random.default_rng(seed)

    def pc_apply(v):
        y = pc.apply_threaded(v).real
        if eps:
            y = y * (1.0 + eps * rng.standard_normal(y.size))
        return y

    b = op.rhs()
    t = time.time()
    # correction: the preconditioned operator as v + P^-1 (A - P) v (pd_set_option "gmres_residual_correction")
    pcm = (lambda v: v + pc_apply(op.delta(v))) if correction else None
    _, its, hist, reason = gmres_lean(op.matvec, pc_apply, b, rtol=rtol, max_it=max_it, pc_matvec=pcm)
    return {"its_at_1e-9": its if reason.startswith("CONVERGED") else None, "hist": hist, "reason": reason,
            "seconds": time.time() - t}


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "mid", "cfg5"]
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        N_x, N_t, gamma = CASES[name]
        rec = {"N_x": N_x, "N_t": N_t, "T": 2.0, "gamma": gamma, "rhs": "manufactured (Build_f/g/IC)"}
        base = run(N_x, N_t, gamma)
        rec["hist"] = base["hist"]
        rec["its"] = {f"{r:g}": count(base["hist"], r) for r in (1e-5, 1e-7, 1e-9)}
        rec["seconds"] = base["seconds"]
        rec["perturbed"] = {}
        for eps in (1e-13, 1e-11):
            p = run(N_x, N_t, gamma, eps=eps)
            rec["perturbed"][f"{eps:g}"] = {f"{r:g}": count(p["hist"], r) for r in (1e-5, 1e-7, 1e-9)}
        corr = run(N_x, N_t, gamma, correction=True)
        rec["correction"] = {"hist": corr["hist"], "its": {f"{r:g}": count(corr["hist"], r) for r in (1e-5, 1e-7, 1e-9)},
                             "perturbed": {}}
        for eps in (1e-13, 1e-11):
            p = run(N_x, N_t, gamma, eps=eps, correction=True)
            rec["correction"]["perturbed"][f"{eps:g}"] = {f"{r:g}": count(p["hist"], r) for r in (1e-5, 1e-7, 1e-9)}
        db[name] = rec
        print(name, rec["its"], rec["perturbed"], "| correction", rec["correction"]["its"],
              rec["correction"]["perturbed"], f"{rec['seconds']:.1f}s", flush=True)
        json.dump(db, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
