"""cfg5 (4096 x 4096) apply time with alpha = 1 and alpha != 1 (Gamma fused into the FFT passes).  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402

for N_x, N_t in ((4096, 4096), (1024, 1024), (300, 81)):
    for al in (1.0, 1e-2):
        with ParaDiagHandle(N_x, N_t, alpha=al) as h:
            x = torch.randn(h.size, dtype=torch.float64, device="cuda:0") + 0j
            y = torch.empty_like(x)
            for _ in range(3):
                h.pc_apply(x, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                h.pc_apply(x, y)
            e1.record()
            torch.cuda.synchronize()
            print(f"N_x={N_x} N_t={N_t} alpha={al:g}: {e0.elapsed_time(e1) / 20:.4f} ms/apply", flush=True)
