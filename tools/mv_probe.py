import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimal_control_paradiag_b200 import ParaDiagHandle
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for (Nx, Nt) in [(4096, 4096), (16384, 4096)]:
    with ParaDiagHandle(Nx, Nt) as h:
        x = torch.randn(h.size, dtype=torch.complex128, device="cuda:0"); y = torch.empty_like(x)
        xr = torch.randn(h.size, dtype=torch.float64, device="cuda:0"); yr = torch.empty_like(xr)
        S = 32 * (Nx + 1) * Nt
        t = timeit(lambda: h.matvec(x, y)); tr = timeit(lambda: h.matvec_real(xr, yr)); tc = timeit(lambda: y.copy_(x))
        print(f"({Nx},{Nt}) matvec {t:.3f} ms = {2*S/t/1e6:.0f} GB/s; real {tr:.3f} ms = {S/tr/1e6:.0f} GB/s; copy {tc:.3f} ms")
