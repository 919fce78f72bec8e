"""Per-kernel times of the single-GPU apply at N_x = 16384, 8192, 4096, 2048 (N_t = 4096): how the streaming passes
scale when the x-range shrinks (what an x-slab of a multi-GPU run sees).  GPU box only."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402

for fuse in ("0", "1"):
    os.environ["PD_FUSE"] = fuse
    for N_x in (16384, 8192, 4096, 2048):
        N_t = 4096
        with ParaDiagHandle(N_x, N_t) as h:
            x = torch.randn(h.size, dtype=torch.float64, device="cuda:0") + 0j
            y = torch.empty_like(x)
            for _ in range(3):
                h.pc_apply(x, y)
            acc = None
            for _ in range(5):
                p = h.pc_apply_profile(x, y)
                acc = p if acc is None else {k: acc[k] + p[k] for k in p}
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                h.pc_apply(x, y)
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps({"fuse": fuse, "N_x": N_x, "ms": e0.elapsed_time(e1) / 20,
                              "kernels_ms": {k: round(v / 5, 4) for k, v in acc.items()}}), flush=True)
        del x, y
        torch.cuda.empty_cache()
