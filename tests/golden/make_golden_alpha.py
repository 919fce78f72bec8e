"""Regenerates tests/golden/alpha_*.npz: the alpha EXTENSION (no upstream counterpart, see oracle/pc_alpha.py).

Produced by the explicit sparse matrix P_alpha factorised with SuperLU (``ExplicitAlphaPC``), the route
that shares no algebra with the decoupled closed form the CUDA kernels use.  "Parity unpinned": the
upstream operator has no alpha.   Run from the repo root:  python tests/golden/make_golden_alpha.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pc_alpha import ExplicitAlphaPC  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(16, 13, 1.0, 0.5), (12, 16, 1.0, 1e-2), (20, 32, 1e-2, 1e-4)]

for (N_x, N_t, gamma, alpha) in CASES:
    rng = np.random.default_rng(0)
    size = 2 * (N_x + 1) * N_t
    x = rng.standard_normal(size) + 1j * rng.standard_normal(size)
    y = ExplicitAlphaPC(N_x, N_t, 2.0, gamma, alpha).apply(x)
    np.savez_compressed(os.path.join(HERE, f"alpha_{N_x}_{N_t}_{gamma:g}_{alpha:g}.npz"),
                        N_x=N_x, N_t=N_t, T=2.0, gamma=gamma, alpha=alpha, x=x, y=y)
print("alpha fixtures written to", HERE)
