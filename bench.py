#!/usr/bin/env python
"""bench.py -- PC applies/s of the ParaDiag block-circulant preconditioner on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one application of the preconditioner (DiagFFTPC.apply, Control_Wave_PC.py:491-553)
to one synthetic complex128 vector (numpy default_rng(0) normal, layout (2, N_x+1, N_t)).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every key.

Workloads (BASELINE.json configs): cfg1 80x81 (upstream default), cfg2 1024x1024,
cfg5 4096x4096, cfg3 16384x4096 (default at N=1: the largest single-GPU configuration and the
one the >=60 %-of-HBM-roofline target is quoted on; its 2.1 GB vectors exceed the 126 MB L2,
so no L2 flush is needed between timed iterations).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": (80, 81), "cfg2": (1024, 1024), "cfg5": (4096, 4096), "cfg3": (16384, 4096),
    "cfg4": (65536, 16384),
}
HUGE = 8 << 30   # vectors above this size are generated on the device and skip the host-side legs
L2_BYTES = 126 * 1024 * 1024


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = upper half of the samples (the sampler also sees idle gaps)
        med = sm[len(sm) * 3 // 4] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(workload, world=1, dist_mode="slab"):
    """The `config` object of the JSON line -- ONE definition shared by both arms (same strings, same keys)."""
    N_x, N_t = WORKLOADS[workload]
    S = 32 * (N_x + 1) * N_t
    return {
        "workload": f"{workload}: 1D wave control N_x={N_x}, N_t={N_t}, T=2, gamma=1, alpha=1",
        "vector_bytes": S, "algorithmic_bytes_per_apply": 6 * S,
        "l2": "inputs larger than L2" if 2 * S > 2 * L2_BYTES else "L2 flushed between timed iterations",
        "parallelism": "1 GPU" if world == 1 else (
            f"x-slabs over {world} GPUs, distributed partition solve (peer stores of 6 N_t values per rank, "
            "no data-path collective)" if dist_mode == "slab" else
            f"space slabs <-> frequency slabs over {world} GPUs (all-to-all)"),
    }


def host_threads():
    """Host threads this process may use: the cpuset it was given (torchrun does not change it)."""
    return len(os.sched_getaffinity(0))


def cpu_reference_apply(N_x, N_t, steps, warmup, threads=None, budget_s=None):
    """The reference's CPU implementation of the path, restated (oracle): scipy.fft (pocketfft, what
    upstream calls at :500-501 / :547-548) with `workers` = all host threads + the fused, pthread-parallel
    per-frequency stage of oracle/csrc/pc_solve.c (rotations + Thomas solves).  Every stage uses all host
    threads; the thread count is set explicitly (OMP_NUM_THREADS etc. of the launcher play no role).
    A step is one full apply; when `budget_s` would be exceeded the per-step sample shrinks to the first
    N_x / f cells (same N_t; every stage is linear in N_x, so seconds scale by f).
    Returns (mean sec per FULL apply, cores, sample description)."""
    import numpy as np
    from oracle import csolve
    from oracle.pc_fast import DiagFFTPCFast
    cores = threads or host_threads()
    csolve.set_num_threads(cores)
    f, pc, x = 1, None, None

    def setup(frac):
        nx = max(8, N_x // frac)
        pcx = DiagFFTPCFast(nx, N_t, 2.0, 1.0, workers=cores)
        rng = np.random.default_rng(0)
        size = 2 * (nx + 1) * N_t
        xx = np.empty(size, dtype=np.complex128)
        step = 1 << 24
        for o in range(0, size, step):
            m = min(step, size - o)
            xx[o:o + m] = rng.standard_normal(m) + 1j * rng.standard_normal(m)
        return pcx, xx

    pc, x = setup(1)
    t = time.perf_counter()
    pc.apply_threaded(x)                       # first touch (page faults, thread start-up): never timed
    first = time.perf_counter() - t
    if budget_s is not None and first * (steps + warmup) > budget_s:
        while f < 64 and first / f * (steps + warmup) > budget_s:
            f *= 2
        pc, x = setup(f)
        pc.apply_threaded(x)
    for _ in range(warmup):
        pc.apply_threaded(x)
    t = time.perf_counter()
    for _ in range(steps):
        pc.apply_threaded(x)
    sec = (time.perf_counter() - t) / steps * (N_x / pc.N_x)
    sample = (f"full {N_x}x{N_t} apply" if f == 1 else
              f"{pc.N_x}x{N_t} apply (1/{f} of the cells, seconds scaled by {N_x / pc.N_x:.3f})")
    sample += f", mean of {steps} after {warmup + 1} warm-up, scipy.fft workers={cores} + threaded C stage"
    return sec, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N_x, N_t = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    # exactly `steps` timed steps after `warmup` warm-ups (plus one untimed first-touch apply); the per-step
    # sample shrinks when the whole run would not finish within a few minutes
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    sec, cores, sample = cpu_reference_apply(N_x, N_t, steps, warmup, budget_s=240.0)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": "pc_applies_per_sec", "value": val, "unit": "applies/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128 (f64 complex)",
        "data": "synthetic", "config": workload_config(args.workload, world, args.dist_mode),
        "cpu_baseline": {"value": val, "unit": "applies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Firedrake/PETSc/MUMPS are not installable here; this is the oracle's CPU restatement "
                "(scipy.fft + fused threaded rotation/Thomas stage) on the box's host cores",
    }
    print(json.dumps(line), flush=True)


def _device_random(torch, size, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(size, dtype=torch.complex128, device=dev)
    xr = torch.view_as_real(x)
    step = 1 << 27
    for o in range(0, size, step):
        xr[o:o + min(step, size - o)].normal_(generator=g)
    return x


def _timed_applies(torch, apply_fn, steps, warmup, flush, barrier):
    """W untimed + exactly K timed applies, CUDA events on the launching stream, barrier + synchronize on both
    sides.  Returns (mean ms per apply on this rank, wall seconds of the timed region)."""
    for _ in range(warmup):
        apply_fn()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t0 = time.perf_counter()
    for s, e in ev:
        if flush is not None:
            flush.zero_()
        s.record()
        apply_fn()
        e.record()
    barrier()
    wall = time.perf_counter() - t0
    return sum(s.elapsed_time(e) for s, e in ev) / steps, wall


def _file_sha16(path):
    import hashlib
    try:
        return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
    except Exception:
        return None


def ncu_traffic(workload, kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this command,
    WITH ITS PROVENANCE (file, sha256, the commit that last touched it) so that a stale capture is visible."""
    fname = {"cfg3": "r02_ncu_full_cfg3.txt"}.get(workload)
    if not fname:
        return None, None
    path = os.path.join(ROOT, "profiles", fname)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_ncu_full_cfg3.txt")
    try:
        blocks = open(path).read().split("== ")
        for blk in blocks:
            if blk.strip() and kernel_key in blk.splitlines()[0]:
                dram = 0.0
                for ln in blk.splitlines():
                    if "dram__bytes_read.sum" in ln or "dram__bytes_write.sum" in ln:
                        val, unit = ln.split()[-2], ln.split()[-1]
                        dram += float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                try:
                    commit = subprocess.run(["git", "log", "-1", "--format=%h", "--", path], cwd=ROOT,
                                            capture_output=True, text=True, timeout=5).stdout.strip() or None
                except Exception:
                    commit = None
                return dram, {"file": os.path.relpath(path, ROOT), "sha256_16": _file_sha16(path), "commit": commit,
                              "note": "read from the committed ncu capture, not measured in this run"}
    except Exception:
        pass
    return None, None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from optimal_control_paradiag_b200 import DiagFFTPC, ParaDiagHandle, petsc_shim

    N_x, N_t = WORKLOADS[args.workload]
    n = N_x + 1
    S = 32 * n * N_t                      # one sweep: both complex128 fields once
    B_pc = 6 * S                          # algorithmic bytes per apply
    peak, peak_src = measured_peak()
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    xn = None
    if world > 1:
        from optimal_control_paradiag_b200.dist import DistributedDiagFFTPC
        handle = DistributedDiagFFTPC(N_x, N_t, device=local, mode=args.dist_mode)
        x = handle.random_local(seed=rank)
        y = torch.empty_like(x)
        apply_fn = lambda: handle.apply(x, y)
    else:
        handle = ParaDiagHandle(N_x, N_t, device=local)
        if S > HUGE:
            x = _device_random(torch, handle.size, dev, 0)
        else:
            rng = np.random.default_rng(0)
            xh = torch.empty(handle.size, dtype=torch.complex128, pin_memory=True)
            xn = xh.numpy()
            step = 1 << 24                # fill in slabs to bound the host memory traffic of the generator
            for o in range(0, handle.size, step):
                m = min(step, handle.size - o)
                xn[o:o + m] = rng.standard_normal(m) + 1j * rng.standard_normal(m)
            x = xh.to(dev)
        y = torch.empty_like(x)
        apply_fn = lambda: handle.pc_apply(x, y)

    flush = None
    if 2 * S <= 2 * L2_BYTES:             # working set may live in L2: flush between iterations
        flush = torch.empty(2 * L2_BYTES // 8, dtype=torch.float64, device=dev)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(warmup):
        apply_fn()
    barrier()
    if sampler:
        sampler.start()
    launches0 = handle.launch_count
    ms, wall = _timed_applies(torch, apply_fn, args.steps, 0, flush, barrier)
    launches = handle.launch_count - launches0
    ms = maxr(ms)
    clocks = sampler.stop() if sampler else None

    line = {
        "metric": "pc_applies_per_sec", "value": 1e3 / ms, "unit": "applies/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128 (f64 complex)", "data": "synthetic",
        "config": workload_config(args.workload, world, args.dist_mode),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "apply_gbs_algorithmic": B_pc / (ms * 1e-3) / 1e9,
        "apply_frac_of_hbm_peak": B_pc / (ms * 1e-3) / 1e9 / peak,
        "wall_s_timed_region": wall,
    }

    if world > 1:
        # ---- per-stage device times at N GPUs (CUDA events inside the library; max over ranks per stage)
        kernels_ms = None
        if getattr(handle, "transport", None) == "peer":
            acc = None
            for _ in range(5):
                barrier()
                p = handle.apply_profile(x, y)
                acc = p if acc is None else {k: acc[k] + p[k] for k in p}
            kernels_ms = {k: maxr(v / 5) for k, v in acc.items()}
        # ---- parity leg: the distributed apply against the single-GPU apply of the same global vector (which
        # tests/test_gpu_fullsize.py checks against the fp64 and the 80-bit oracle at this size) on rank 0
        parity = None
        try:
            handle.apply(x, y)
            xg = handle.gather_to_global(x)
            yg = handle.gather_to_global(y)
            pr = [0.0, 0.0, 0.0]
            if rank == 0:
                with ParaDiagHandle(N_x, N_t, device=local) as h1:
                    ref = h1.pc_apply(xg)
                    pr[0] = float(torch.linalg.norm(yg - ref) / torch.linalg.norm(ref))
                    pr[1] = float(yg.view(2, n, N_t)[:, [0, -1], :].abs().max())
                    pr[2] = float(torch.linalg.norm(ref))
                    del ref
            del xg, yg
            torch.cuda.empty_cache()
            timed_out = False
            if getattr(handle, "transport", None) == "peer":
                timed_out = handle.backend.slab_comm_status()[0]
            parity = {"rel_err": pr[0], "boundary_zero": pr[1] == 0.0, "exchange_timed_out": bool(maxr(timed_out)),
                      "against": "single-GPU pd_pc_apply of the gathered global vector, on rank 0",
                      "tolerance": 1e-10, "ok": bool(pr[0] < 1e-10 and pr[1] == 0.0)}
        except Exception as ex:  # pragma: no cover
            parity = {"error": str(ex)[:300]}

        # ---- end to end at N GPUs: every rank's node-slab block lives in pinned HOST memory (a PETSc Vec of a
        # spatial decomposition); each step = H2D of the block, distributed apply, D2H of the result
        xh = torch.empty(handle.local_size, dtype=torch.complex128, pin_memory=True)
        yh = torch.empty(handle.local_size, dtype=torch.complex128, pin_memory=True)
        xh.copy_(x)
        e2e_steps = max(1, min(args.steps, 10))
        handle.apply_host(xh, yh)
        barrier()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            handle.apply_host(xh, yh)
        barrier()
        te = maxr((time.perf_counter() - t1) / e2e_steps * 1e3)
        nb = torch.tensor([float(handle.local_size * 16)], dtype=torch.float64, device=dev)
        dist.all_reduce(nb, op=dist.ReduceOp.SUM)
        e2e_dist = {"value": 1e3 / te, "unit": "applies/s", "h2d_bytes_per_step": int(nb.item()),
                    "d2h_bytes_per_step": int(nb.item()), "ms_per_step": te, "steps": e2e_steps,
                    "api": "DistributedDiagFFTPC.apply_host(x, y): every rank's node-slab block in pinned host memory"}
        del xh, yh

        real_dist = None
        if args.dist_mode == "slab":
            # the same apply on float64 blocks (the real problem): half spectrum through the slab-distributed solve
            try:
                xr = torch.randn(handle.local_size, dtype=torch.float64, device=dev)
                yr = torch.empty_like(xr)
                tr, _ = _timed_applies(torch, lambda: handle.apply_real(xr, yr), 10, 3, None, barrier)
                tr = maxr(tr)
                real_dist = {"ms_per_step": tr, "applies_per_sec": 1e3 / tr,
                             "note": "DistributedDiagFFTPC.apply_real on float64 node-slab blocks (half spectrum)"}
                del xr, yr
            except Exception as ex:  # unsupported N_t
                real_dist = {"error": str(ex)[:300]}

        gmres_dist = None
        # (a restart-300 Krylov basis of multi-GB local vectors does not fit: skip the solve leg there)
        if args.dist_mode == "slab" and not args.no_gmres and handle.local_size * 16 <= (4 << 30):
            try:
                b = handle.build_rhs()
                handle.gmres(b, rtol=1e-7)
                barrier()
                t1 = time.perf_counter()
                _, its, hist, reason = handle.gmres(b, rtol=1e-7)
                barrier()
                gmres_dist = {"seconds": time.perf_counter() - t1, "iterations": its, "reason": reason, "rtol": 1e-7,
                              "rhs": "manufactured (Build_f/g/IC)"}
                del b
            except Exception as ex:  # pragma: no cover
                gmres_dist = {"error": str(ex)[:300]}
        line["e2e"] = e2e_dist
        line["dist"] = handle.describe()
        line["kernels_ms"] = kernels_ms
        line["parity"] = parity
        line["gmres"] = gmres_dist
        line["real_input_apply"] = real_dist
    else:
        # per-kernel device times (CUDA events on the launching stream), live
        prof = None
        reps = 5
        for _ in range(reps):
            p = handle.pc_apply_profile(x, y)
            prof = p if prof is None else {k: prof[k] + p[k] for k in p}
        prof = {k: v / reps for k, v in prof.items()}
        tot = sum(prof.values())
        # algorithmic bytes per launch: each FFT pass and the solve pass read S and write S;
        # the solve pass is pass A + interface + pass B, of which pass B carries the read+write sweep
        dom = max(prof, key=prof.get)
        alg = {"ifft": 2 * S, "fft": 2 * S, "passB": 2 * S, "passA": S, "interface": 0.3 * S}[dom]
        fftk = "pd_fft_16k_l2_kernel" if N_t == 16384 else ("pd_fft_pow2_kernel" if (N_t & (N_t - 1)) == 0 and N_t >= 64
                                                         else "pd_fft_generic_kernel")
        names = {"ifft": fftk + "<inv>", "fft": fftk + "<fwd>", "passA": "pd_solve_passA_kernel",
                 "interface": "pd_solve_iface_thomas_kernel", "passB": "pd_solve_passB_kernel"}
        ach = alg / (prof[dom] * 1e-3) / 1e9
        key = {"ifft": "pd_fft_pow2_kernel<16, 16, 16, 1, 1>", "fft": "pd_fft_pow2_kernel<16, 16, 16, 1, 0>",
               "passA": "pd_solve_passA_kernel", "passB": "pd_solve_passB_kernel",
               "interface": "pd_solve_iface_thomas_kernel"}[dom]
        traffic, prov = ncu_traffic(args.workload, key)
        line["roofline"] = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": traffic, "traffic_provenance": prov,
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                            "ms_per_launch": prof[dom], "share_of_apply": prof[dom] / tot}
        line["kernels_ms"] = prof
        line["roofline_whole_apply"] = {"achieved": B_pc / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": B_pc / (ms * 1e-3) / 1e9 / peak,
                                        "frac_of_nominal_8000": B_pc / (ms * 1e-3) / 1e9 / 8000.0}

        if xn is None:
            line["e2e"] = None
            line["note"] = "vectors > 8 GiB: generated on the device, host-side legs (e2e, cpu_baseline, gmres) skipped"
        else:
            # end to end through the reference-facing python PC with HOST vectors (pinned), every step:
            # H2D of x, apply, D2H of y
            DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=1.0, device=local)
            pc = petsc_shim.PC()
            pc.setPythonContext(DiagFFTPC())
            pc.setUp()
            yh = torch.empty(handle.size, dtype=torch.complex128, pin_memory=True)
            xv, yv = petsc_shim.Vec(xn), petsc_shim.Vec(yh.numpy())
            xv._a, yv._a = xn, yh.numpy()
            e2e_steps = max(1, min(args.steps, 10))
            pc.apply(xv, yv)
            t1 = time.perf_counter()
            for _ in range(e2e_steps):
                pc.apply(xv, yv)
            e2e_ms = (time.perf_counter() - t1) / e2e_steps * 1e3
            line["e2e"] = {"value": 1e3 / e2e_ms, "unit": "applies/s", "h2d_bytes_per_step": S, "d2h_bytes_per_step": S,
                           "ms_per_step": e2e_ms, "steps": e2e_steps,
                           "api": "DiagFFTPC.apply(pc, x, y) with host Vec buffers (pinned)"}
            # the same with PAGEABLE host Vecs (what PETSc allocates): the library page-locks each buffer once
            # (cudaHostRegister cache); the first call pays the registration and is reported separately
            try:
                xp = np.empty_like(xn)
                xp[...] = xn
                yp = np.empty_like(xn)
                xv2, yv2 = petsc_shim.Vec(xp), petsc_shim.Vec(yp)
                xv2._a, yv2._a = xp, yp
                pc.apply(xv2, yv2)
                t1 = time.perf_counter()
                for _ in range(max(1, e2e_steps // 2)):
                    pc.apply(xv2, yv2)
                plain_ms = (time.perf_counter() - t1) / max(1, e2e_steps // 2) * 1e3
                pc.getPythonContext().handle.set_option("host_register", 1)     # diagfft_register_vecs
                t1 = time.perf_counter()
                pc.apply(xv2, yv2)
                first_ms = (time.perf_counter() - t1) * 1e3
                t1 = time.perf_counter()
                for _ in range(e2e_steps):
                    pc.apply(xv2, yv2)
                pg_ms = (time.perf_counter() - t1) / e2e_steps * 1e3
                pc.getPythonContext().handle.host_unregister_all()
                line["e2e_pageable"] = {"value": 1e3 / pg_ms, "unit": "applies/s", "ms_per_step": pg_ms,
                                        "registration_call_ms": first_ms, "steps": e2e_steps,
                                        "ms_per_step_unregistered": plain_ms,
                                        "api": "DiagFFTPC.apply(pc, x, y) with pageable numpy buffers; option "
                                               "diagfft_register_vecs: page-locked once by pd_pc_apply_host"}
                del xp, yp, xv2, yv2
            except Exception as ex:  # pragma: no cover
                line["e2e_pageable"] = {"error": str(ex)[:300]}
            # the same call with REAL host Vecs (a real-scalar PETSc build): float64 arrays, half-spectrum path,
            # half the PCIe bytes.  Side record: the headline e2e above keeps the reference's complex Vecs.
            try:
                if handle.real_path_supported:
                    xr = torch.empty(handle.size, dtype=torch.float64, pin_memory=True)
                    xr.numpy()[...] = xn.real
                    yr = torch.empty(handle.size, dtype=torch.float64, pin_memory=True)
                    pc.apply(xr.numpy(), yr.numpy())
                    t1 = time.perf_counter()
                    for _ in range(e2e_steps):
                        pc.apply(xr.numpy(), yr.numpy())
                    r_ms = (time.perf_counter() - t1) / e2e_steps * 1e3
                    line["e2e_real"] = {"value": 1e3 / r_ms, "unit": "applies/s", "ms_per_step": r_ms,
                                        "h2d_bytes_per_step": S // 2, "d2h_bytes_per_step": S // 2, "steps": e2e_steps,
                                        "api": "DiagFFTPC.apply(pc, x, y) with float64 host buffers (pinned): "
                                               "pd_pc_apply_real_host"}
                    del xr, yr
            except Exception as ex:  # pragma: no cover
                line["e2e_real"] = {"error": str(ex)[:300]}
            pc.destroy()
            DiagFFTPC._defaults = {}

            # GMRES time-to-solution on the reference's manufactured problem (secondary metric)
            if not args.no_gmres:
                line.update(_gmres_legs(torch, handle, dev))

            # CPU baseline: the oracle's restatement on the host cores, bounded sample
            if not args.no_cpu:
                try:
                    sec, cores, sample = cpu_reference_apply(N_x, N_t, steps=3, warmup=0, budget_s=25.0)
                    line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "applies/s", "cores": cores, "kind": "port",
                                            "sample": sample}
                except Exception as ex:  # pragma: no cover
                    line["cpu_baseline"] = {"value": None, "unit": "applies/s", "cores": 0, "kind": "port",
                                            "sample": f"failed: {ex}"}

    # ---- sub-record: the scaling configuration of BASELINE (cfg4, 65536 x 16384, 34 GB vectors) at this N
    if args.workload == "cfg3" and not args.no_cfg4:
        del x, y, apply_fn
        if world > 1:
            del handle
        else:
            handle.close()
        torch.cuda.empty_cache()
        try:
            line["cfg4"] = _cfg4_record(torch, dist, world, rank, local, dev, args, peak, barrier, maxr)
        except Exception as ex:  # pragma: no cover
            line["cfg4"] = {"error": str(ex)[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _gmres_legs(torch, handle, dev):
    out = {}
    try:
        b = handle.build_rhs()
        handle.gmres(b, rtol=1e-7)           # warm-up: allocates the Krylov basis
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        _, its, hist, reason = handle.gmres(b, rtol=1e-7)
        torch.cuda.synchronize()
        out["gmres"] = {"seconds": time.perf_counter() - t1, "iterations": its, "reason": reason,
                        "rtol": 1e-7, "rhs": "manufactured (Build_f/g/IC)"}
        del b
        # the same solve on float64 vectors (real problem): half-spectrum PC, half the BLAS-1 bytes
        try:
            br = handle.build_rhs_real()
            handle.gmres_real(br, rtol=1e-7)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            _, its, hist, reason = handle.gmres_real(br, rtol=1e-7)
            torch.cuda.synchronize()
            out["gmres_real_vectors"] = {"seconds": time.perf_counter() - t1, "iterations": its,
                                         "reason": reason, "rtol": 1e-7}
            xr = torch.randn(handle.size, dtype=torch.float64, device=dev)
            yr = torch.empty_like(xr)
            for _ in range(3):
                handle.pc_apply_real(xr, yr)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                handle.pc_apply_real(xr, yr)
            e1.record()
            torch.cuda.synchronize()
            out["real_input_apply"] = {"ms_per_step": e0.elapsed_time(e1) / 10,
                                       "applies_per_sec": 1e4 / e0.elapsed_time(e1),
                                       "note": "pd_pc_apply_real on float64 vectors (half spectrum)"}
            del br, xr, yr
        except Exception as ex:  # unsupported N_t etc.
            out["gmres_real_vectors"] = {"error": str(ex)[:300]}
    except Exception as ex:  # pragma: no cover
        out["gmres"] = {"error": str(ex)[:300]}
    return out


def _cfg4_record(torch, dist, world, rank, local, dev, args, peak, barrier, maxr):
    """PC applies/s at cfg4 (N_x = 65536, N_t = 16384) on the same N GPUs, device-resident, plus a parity check
    that needs no second copy of the 34 GB problem: the normwise backward error ||P y - x|| / (||P|| ||y||) of the
    output on interior rows, P the explicit block-circulant stencil (pd_pc_matvec, itself checked against the
    explicit sparse matrix of the oracle at small sizes), evaluated on rank 0 after gathering x and y."""
    from optimal_control_paradiag_b200 import ParaDiagHandle
    N_x, N_t = WORKLOADS["cfg4"]
    n = N_x + 1
    S = 32 * n * N_t
    steps = max(3, min(args.steps, 10))
    if world > 1:
        from optimal_control_paradiag_b200.dist import DistributedDiagFFTPC
        h4 = DistributedDiagFFTPC(N_x, N_t, device=local, mode=args.dist_mode)
        x = _device_random(torch, h4.local_size, dev, 100 + rank)
        y = torch.empty_like(x)
        fn = lambda: h4.apply(x, y)
    else:
        h4 = ParaDiagHandle(N_x, N_t, device=local)
        x = _device_random(torch, h4.size, dev, 100)
        y = torch.empty_like(x)
        fn = lambda: h4.pc_apply(x, y)
    ms, _ = _timed_applies(torch, fn, steps, 3, None, barrier)
    ms = maxr(ms)
    rec = {"workload": workload_config("cfg4", world, args.dist_mode)["workload"], "ms_per_step": ms,
           "value": 1e3 / ms, "unit": "applies/s", "steps": steps, "warmup": 3, "vector_bytes": S,
           "apply_gbs_algorithmic": 6 * S / (ms * 1e-3) / 1e9, "apply_frac_of_hbm_peak": 6 * S / (ms * 1e-3) / 1e9 / peak}
    if world > 1 and getattr(h4, "transport", None) == "peer":
        acc = None
        for _ in range(3):
            barrier()
            p = h4.apply_profile(x, y)
            acc = p if acc is None else {k: acc[k] + p[k] for k in p}
        rec["kernels_ms"] = {k: maxr(v / 3) for k, v in acc.items()}
        rec["transport"] = h4.transport
    elif world == 1:
        p = h4.pc_apply_profile(x, y)
        rec["kernels_ms"] = p
    # parity: backward error on rank 0
    try:
        fn()
        if world > 1:
            ncount, noff = h4.ncount, h4.noff
            timed_out = h4.backend.slab_comm_status()[0] if getattr(h4, "transport", None) == "peer" else False
            del h4
            torch.cuda.empty_cache()
            if rank == 0:
                xg = torch.empty(2 * n * N_t, dtype=torch.complex128, device=dev)
                yg = torch.empty(2 * n * N_t, dtype=torch.complex128, device=dev)
            for src, buf in ((x, "x"), (y, "y")):
                full = (xg if buf == "x" else yg) if rank == 0 else None
                for r in range(world):
                    if r == 0:
                        if rank == 0:
                            full.view(2, n, N_t)[:, noff[0]:noff[1], :] = src.view(2, ncount[0], N_t)
                        continue
                    if rank == r:
                        dist.send(torch.view_as_real(src), dst=0)
                    elif rank == 0:
                        tmp = torch.empty(2 * ncount[r] * N_t, dtype=torch.complex128, device=dev)
                        dist.recv(torch.view_as_real(tmp), src=r)
                        full.view(2, n, N_t)[:, noff[r]:noff[r + 1], :] = tmp.view(2, ncount[r], N_t)
                        del tmp
            del x, y
            torch.cuda.empty_cache()
        else:
            xg, yg, timed_out = x, y, False
            h4.close()
            torch.cuda.empty_cache()
        vals = [0.0, 0.0]
        if rank == 0:
            with ParaDiagHandle(N_x, N_t, device=local) as hp:
                r = hp.pc_matvec(yg)
                r.sub_(xg)
                R = r.view(2, n, N_t)[:, 1:-1, :]
                hh, dt = 1.0 / N_x, 2.0 / N_t
                normP = 4 * hh + 4 * dt * dt / hh + dt * dt * hh
                vals[0] = float(torch.linalg.norm(R) / (normP * torch.linalg.norm(yg)))
                vals[1] = float(yg.view(2, n, N_t)[:, [0, -1], :].abs().max())
                del r
            del xg, yg
        torch.cuda.empty_cache()
        rec["parity"] = {"backward_error": vals[0], "boundary_zero": vals[1] == 0.0,
                         "exchange_timed_out": bool(maxr(timed_out)), "tolerance": 1e-13,
                         "against": "||P y - x|| / (||P|| ||y||) on interior rows, explicit block-circulant stencil, rank 0",
                         "ok": bool(vals[0] < 1e-13 and vals[1] == 0.0)}
    except Exception as ex:  # pragma: no cover
        rec["parity"] = {"error": str(ex)[:300]}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--dist-mode", default="slab", choices=["slab", "alltoall"],
                    help="multi-GPU decomposition of the solve stage (see dist.py)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gmres", action="store_true", help="skip the GMRES time-to-solution leg")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the cfg4 (65536 x 16384) sub-record")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
