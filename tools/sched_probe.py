"""Times the single-GPU apply under the plain and the node-slab interleaved schedule (PD_SCHED*, read by
pd_create) and checks that both give the same bits.  GPU box only:  python tools/sched_probe.py [cfg3|cfg5] ..."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402

SIZES = {"cfg3": (16384, 4096), "cfg5": (4096, 4096), "cfg3h": (16384, 2048), "big": (32768, 8192)}


def timed(h, x, y, reps=20):
    for _ in range(3):
        h.pc_apply(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        h.pc_apply(x, y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timed_graph(h, x, y, reps=20):
    """The same with the apply captured once in a CUDA graph and replayed: no host launch cost."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            h.pc_apply(x, y)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        h.pc_apply(x, y)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    names = sys.argv[1:] or ["cfg3"]
    for name in names:
        N_x, N_t = SIZES[name]
        g = torch.Generator(device="cuda:0").manual_seed(0)
        size = 2 * (N_x + 1) * N_t
        x = torch.randn(size, dtype=torch.float64, device="cuda:0", generator=g) + 0j
        x = x + 1j * torch.randn(size, dtype=torch.float64, device="cuda:0", generator=g)
        y = torch.empty_like(x)
        os.environ["PD_SCHED"] = "plain"
        with ParaDiagHandle(N_x, N_t) as h:
            ms0 = timed(h, x, y)
            ref = y.clone()
            msg0 = timed_graph(h, x, y)
        print(json.dumps({"size": name, "sched": "plain", "ms": ms0, "ms_graph": msg0}), flush=True)
        os.environ["PD_SCHED"] = "interleave"
        for streams in (1, 2):
            for chunks in (6, 13, 26, 52, 104):
                os.environ["PD_SCHED_STREAMS"] = str(streams)
                os.environ["PD_SCHED_CHUNKS"] = str(chunks)
                with ParaDiagHandle(N_x, N_t) as h:
                    ms = timed(h, x, y)
                    same = bool(torch.equal(y, ref))
                    try:
                        msg = timed_graph(h, x, y)
                        same = same and bool(torch.equal(y, ref))
                    except Exception as ex:
                        msg = str(ex)[:200]
                print(json.dumps({"size": name, "sched": "interleave", "streams": streams, "chunks": chunks,
                                  "slab_MB": chunks * 17 * 2 * N_t * 16 / 2**20, "ms": ms, "ms_graph": msg, "same_bits": same,
                                  "speedup": ms0 / ms}), flush=True)
        del x, y, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
