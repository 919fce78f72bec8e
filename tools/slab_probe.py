#!/usr/bin/env python
"""All G x-slab ranks of one distributed apply on ONE GPU (LocalSlabGroup): the aggregate per-rank kernel work of the
slab mode without any link or skew effect, against the single-GPU apply of the same vector.  The difference divided by
G is what every rank pays for being a slab (short grids, interface / separator kernels).
Usage: python tools/slab_probe.py [N_x N_t G [reps]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402
from optimal_control_paradiag_b200.dist import LocalSlabGroup  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    N_x, N_t, G = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (16384, 4096, 8)))
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    real = os.environ.get("PROBE_REAL") == "1"
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, G) as grp:
        g0 = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(h.size, dtype=torch.float64, device="cuda", generator=g0)
        if not real:
            x = x.to(torch.complex128)
        y = torch.empty_like(x)
        one = timed(lambda: (h.pc_apply_real(x, y) if real else h.pc_apply(x, y)), reps)
        xs = grp.scatter(x)
        agg = timed(lambda: grp.apply_blocks(xs, real=real), reps)
        print(f"{N_x}x{N_t} G={G} real={real}: single-GPU apply {one:.4f} ms | all {G} slab ranks on one GPU {agg:.4f} ms "
              f"| per rank {agg / G:.4f} ms (ideal {one / G:.4f}), slab overhead per rank {(agg - one) / G * 1e3:.1f} us")


if __name__ == "__main__":
    main()
