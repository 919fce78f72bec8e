import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast
for (Nx, Nt) in [(4096, 8), (9600, 8), (9536, 8), (16384, 8), (16384, 12), (20000, 5)]:
    rng = np.random.default_rng(0)
    size = 2 * (Nx + 1) * Nt
    x = rng.standard_normal(size) + 1j * rng.standard_normal(size)
    ref = DiagFFTPCFast(Nx, Nt, dtype=np.longdouble).apply(x).reshape(2, Nx + 1, Nt)
    with ParaDiagHandle(Nx, Nt) as h:
        y = h.pc_apply_host(x).reshape(2, Nx + 1, Nt)
    d = y - ref
    pk = np.sqrt((np.abs(d) ** 2).sum(axis=(0, 1))) / np.sqrt((np.abs(ref) ** 2).sum(axis=(0, 1)))
    rows = [Nx - 1]
    while rows[-1] > 32: rows.append(rows[-1] // 17)
    print(Nx, Nt, "levels", rows, "per-k err", ["%.1e" % float(v) for v in pk])
