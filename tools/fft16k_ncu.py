"""A few launches of the N_t = 16384 FFT kernels (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimal_control_paradiag_b200 import ParaDiagHandle
Nx, N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 16384
with ParaDiagHandle(Nx, N) as h:
    x = torch.randn(h.size, dtype=torch.complex128, device="cuda:0")
    w = torch.empty_like(x)
    for _ in range(3):
        h.stage_fft(x, w, 2 * (Nx + 1), True)
        h.stage_fft(w, x, 2 * (Nx + 1), False)
    torch.cuda.synchronize()
print("ok")
