"""Times the single-GPU apply with the fused inverse-FFT + pass-A kernel (PD_FUSE*, read by pd_create) against the
two separate kernels and checks that both give the same bits.  GPU box only:  python tools/fuse_probe.py [cfg3] ..."""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402

SIZES = {"cfg3": (16384, 4096), "cfg5": (4096, 4096), "cfg2": (1024, 1024), "n2048": (16384, 2048),
         "n8192": (8192, 8192), "odd": (5000, 1024)}


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


def main():
    names = [a for a in sys.argv[1:] if a in SIZES] or ["cfg3"]
    quick = "quick" in sys.argv
    for name in names:
        N_x, N_t = SIZES[name]
        g = torch.Generator(device="cuda:0").manual_seed(0)
        size = 2 * (N_x + 1) * N_t
        x = torch.randn(size, dtype=torch.float64, device="cuda:0", generator=g) + 0j
        x = x + 1j * torch.randn(size, dtype=torch.float64, device="cuda:0", generator=g)
        y = torch.empty_like(x)
        os.environ["PD_FUSE"] = "0"
        with ParaDiagHandle(N_x, N_t) as h:
            ms0 = timed(h, x, y)
            ref = y.clone()
            prof0 = h.pc_apply_profile(x, y)
        print(json.dumps({"size": name, "fuse": 0, "ms": ms0, "kernels_ms": prof0}), flush=True)
        os.environ["PD_FUSE"] = "1"
        combos = [(8, 2, 1)] if quick else list(itertools.product((4, 8, 16, 32), (1, 2, 4), (1, 2)))
        for chunks, cpb, lag in combos:
            if cpb > chunks:
                continue
            os.environ["PD_FUSE_CHUNKS"], os.environ["PD_FUSE_CPB"], os.environ["PD_FUSE_LAG"] = str(chunks), str(cpb), str(lag)
            with ParaDiagHandle(N_x, N_t) as h:
                y.zero_()
                ms = timed(h, x, y)
                same = bool(torch.equal(y, ref))
                prof = h.pc_apply_profile(x, y)
            print(json.dumps({"size": name, "fuse": 1, "chunks": chunks, "cpb": cpb, "lag": lag, "ms": ms,
                              "same_bits": same, "speedup": ms0 / ms, "fused_ms": prof["ifft"]}), flush=True)
        del x, y, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
