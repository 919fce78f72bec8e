import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from optimal_control_paradiag_b200 import ParaDiagHandle
from oracle.pc_fast import DiagFFTPCFast
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for (Nx, Nt) in [(16, 128), (33, 256), (100, 512), (64, 1024), (20, 2048), (17, 4096), (9, 8192), (8, 16384), (1024, 1024)]:
    with ParaDiagHandle(Nx, Nt) as h:
        x = np.random.default_rng(1).standard_normal(h.size)
        ref = DiagFFTPCFast(Nx, Nt).apply(x + 0j)
        y = h.pc_apply_real(torch.tensor(x, device="cuda:0")).cpu().numpy()
        yc = h.pc_apply(torch.tensor(x + 0j, device="cuda:0")).cpu().numpy()
        print(f"({Nx},{Nt}) real path vs oracle {np.linalg.norm(y - ref.real)/np.linalg.norm(ref):.2e}  vs complex path {np.linalg.norm(y - yc.real)/np.linalg.norm(yc):.2e}", flush=True)
for (Nx, Nt) in [(4096, 4096), (16384, 4096)]:
    with ParaDiagHandle(Nx, Nt) as h:
        xr = torch.randn(h.size, dtype=torch.float64, device="cuda:0"); yr = torch.empty_like(xr)
        xc = xr.to(torch.complex128); yc = torch.empty_like(xc)
        tr = timeit(lambda: h.pc_apply_real(xr, yr)); tc = timeit(lambda: h.pc_apply(xc, yc))
        d = float(torch.linalg.norm(yr - yc.real) / torch.linalg.norm(yc.real))
        print(f"({Nx},{Nt}) real {tr:.3f} ms  complex {tc:.3f} ms  ratio {tc/tr:.2f}  diff {d:.2e}")

Nx, Nt = 16384, 4096
with ParaDiagHandle(Nx, Nt) as h:
    n = Nx + 1; K = (Nt // 2 + 1 + 7) // 8 * 8
    xr = torch.randn(h.size, dtype=torch.float64, device="cuda:0"); yr = torch.empty_like(xr)
    w = torch.empty(2 * n * K, dtype=torch.complex128, device="cuda:0")
    t1 = timeit(lambda: h.stage_rfft(xr, w, 2 * n, True))
    t2 = timeit(lambda: h.stage_solve_half(w))
    t3 = timeit(lambda: h.stage_rfft(w, yr, 2 * n, False))
    print(f"real stages: r2c {t1:.3f} solve {t2:.3f} c2r {t3:.3f} ms")
