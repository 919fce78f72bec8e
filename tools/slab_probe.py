"""Developer probe: per-stage device time of one rank's slab-mode work (no communication)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimal_control_paradiag_b200 import ParaDiagHandle
def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
Nx, Nt = 16384, 4096
for G in (2, 4, 8):
    r = G // 2
    h = ParaDiagHandle(Nx, Nt, slab_rank=r, slab_count=G)
    n_r = (Nx + 1) // G + (1 if r < (Nx + 1) % G else 0)
    x = torch.randn(2 * n_r * Nt, dtype=torch.complex128, device="cuda:0")
    w = torch.empty_like(x); y = torch.empty_like(x)
    out = torch.empty(6 * Nt, dtype=torch.complex128, device="cuda:0")
    gathered = torch.randn(G * 6 * Nt, dtype=torch.complex128, device="cuda:0")
    t1 = timeit(lambda: h.stage_fft(x, w, 2 * n_r, True))
    t2 = timeit(lambda: h.slab_reduce(w, out))
    t3 = timeit(lambda: h.slab_finish(w, gathered))
    t4 = timeit(lambda: h.stage_fft(w, y, 2 * n_r, False))
    def full():
        h.stage_fft(x, w, 2 * n_r, True); h.slab_reduce(w, out); h.slab_finish(w, gathered); h.stage_fft(w, y, 2 * n_r, False)
    t5 = timeit(full)
    print(f"G={G}: ifft {t1:.0f} us, reduce {t2:.0f} us, finish {t3:.0f} us, fft {t4:.0f} us, sum {t1+t2+t3+t4:.0f}, back-to-back {t5:.0f} us (ideal 1-GPU/G = {2720/G:.0f})")
    h.close()
