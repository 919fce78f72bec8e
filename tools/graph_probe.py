#!/usr/bin/env python
"""Launch-bound sizes: one apply as five stream launches against one CUDA-graph replay (L2 flushed between timed
applies, CUDA events around the apply only).  Usage: python tools/graph_probe.py [N_x N_t ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402


def timed(fn, flush, steps=60, warm=5):
    for _ in range(warm):
        flush()
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ts) / len(ts), ts[len(ts) // 2]


def main():
    sizes = [int(v) for v in sys.argv[1:]] or [80, 81, 1024, 1024, 2048, 4096]
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for N_x, N_t in zip(sizes[::2], sizes[1::2]):
        with ParaDiagHandle(N_x, N_t) as h:
            g0 = torch.Generator(device="cuda").manual_seed(0)
            x = torch.randn(h.size, dtype=torch.float64, device="cuda", generator=g0).to(torch.complex128)
            y = torch.empty_like(x)
            small = 2 * x.numel() * 16 <= (252 << 20)
            flush = (lambda: junk.fill_(1)) if small else (lambda: None)
            h.pc_apply(x, y)
            direct = timed(lambda: h.pc_apply(x, y), flush)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                h.pc_apply(x, y)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                h.pc_apply(x, y)
            graph = timed(g.replay, flush)
            print(f"{N_x}x{N_t}: stream launches mean {direct[0] * 1e3:.1f} us median {direct[1] * 1e3:.1f} us | "
                  f"graph replay mean {graph[0] * 1e3:.1f} us median {graph[1] * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
