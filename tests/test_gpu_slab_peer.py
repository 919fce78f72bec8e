"""The peer-store exchange of the slab-distributed apply (pd_slab_apply) on real devices.

* one GPU, one process, G slab handles (`LocalSlabGroup`): the whole distributed code path -- local elimination,
  functionals stored into every rank's exchange buffer, flags / epochs / parities, separator solve, back
  substitution -- against the single-GPU apply.  Runs wherever ONE GPU is leased.
* two GPUs, one process: one slab per device through plain peer access (and handles on two devices in one
  process, which every entry point must survive without disturbing the caller's current device).
* >= 2 GPUs, one process per GPU: cudaIpc-mapped buffers, NVLink peer stores, ranks deliberately skewed.
"""
import os
import socket
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from optimal_control_paradiag_b200 import ParaDiagHandle  # noqa: E402
from optimal_control_paradiag_b200.dist import LocalSlabGroup  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rand_global(size, seed=0, real=False):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(size, dtype=torch.float64, device=DEV, generator=g)
    if real:
        return x
    return x + 1j * torch.randn(size, dtype=torch.float64, device=DEV, generator=g)


def relerr(a, b):
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


@pytest.mark.parametrize("N_x,N_t,G", [(80, 81, 2), (255, 128, 3), (1024, 1024, 4), (1024, 1024, 8), (37, 16, 2),
                                        (4096, 64, 8), (40, 16384, 4), (2047, 256, 5), (16384, 128, 8)])
def test_peer_exchange_on_one_gpu_equals_single_gpu_apply(N_x, N_t, G):
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, G) as grp:
        for rep in range(5):                                   # epochs 1..5: both parities, flags reused
            x = rand_global(h.size, seed=rep)
            ref = h.pc_apply(x)
            y = grp.apply(x)
            assert relerr(y, ref) < 1e-10, (rep, relerr(y, ref))
            Y = y.view(2, N_x + 1, N_t)
            assert float(Y[:, [0, -1], :].abs().max()) == 0.0
        st = grp.status()
        assert all(not to for to, _ in st) and all(ep == 5 for _, ep in st), st


@pytest.mark.parametrize("N_x,N_t,G", [(255, 128, 3), (1024, 1024, 4), (40, 16384, 2), (4096, 256, 8),
                                        (80, 81, 2), (255, 100, 3), (64, 16, 4)])   # + the shared-memory pair kernel
def test_peer_exchange_real_input_path_on_one_gpu(N_x, N_t, G):
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, G) as grp:
        for rep in range(3):
            x = rand_global(h.size, seed=10 + rep, real=True)
            ref = h.pc_apply_real(x)
            y = grp.apply(x, real=True)
            assert relerr(y, ref) < 1e-10
        # complex and real applies interleave on the same exchange buffers (epochs keep counting)
        xc = rand_global(h.size, seed=99)
        assert relerr(grp.apply(xc), h.pc_apply(xc)) < 1e-10
        assert all(not to and ep == 4 for to, ep in grp.status())


@pytest.mark.parametrize("N_x,N_t,G,real", [(1024, 1024, 4, False), (4096, 4096, 8, False), (2047, 2048, 3, True),
                                             (300, 16384, 2, False), (4096, 4096, 8, True)])
def test_frequency_halves_of_the_slab_apply(N_x, N_t, G, real):
    # pd_slab_apply runs its per-frequency stage as two frequency halves (on two streams when every rank has its own
    # GPU); here the halves run one after the other on one stream: same kernels, same column ranges
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, G, split_on_one_stream=True) as grp:
        for rep in range(3):
            x = rand_global(h.size, seed=rep, real=real)
            ref = h.pc_apply_real(x) if real else h.pc_apply(x)
            assert relerr(grp.apply(x, real=real), ref) < 1e-10
        assert all(not to and ep == 3 for to, ep in grp.status())


def test_peer_exchange_at_cfg3_with_eight_slabs_on_one_gpu():
    # the driver's scaling configuration (cfg3, 8 ranks) with all eight slabs on the one leased GPU
    N_x, N_t, G = 16384, 4096, 8
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, G) as grp:
        x = rand_global(h.size)
        ref = h.pc_apply(x)
        y = grp.apply(x)
        assert relerr(y, ref) < 1e-10
        assert all(not to for to, _ in grp.status())


def test_an_undelivered_peer_times_out_instead_of_hanging():
    # only rank 0 runs: its separator kernel must give up after the bounded wait and raise the status flag
    N_x, N_t, G = 64, 128, 2
    with LocalSlabGroup(N_x, N_t, G) as grp:
        xs = grp.scatter(rand_global(2 * (N_x + 1) * N_t))
        y0 = torch.empty_like(xs[0])
        grp.handles[0].slab_apply_begin(xs[0])
        t = time.time()
        grp.handles[0].slab_apply_end(y0)
        timed_out, epoch = grp.handles[0].slab_comm_status()
        assert timed_out and epoch == 1 and time.time() - t < 60.0


def test_handles_on_two_devices_in_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N_x, N_t = 300, 128
    torch.cuda.set_device(0)
    with ParaDiagHandle(N_x, N_t, device=0) as h0, ParaDiagHandle(N_x, N_t, device=1) as h1:
        assert torch.cuda.current_device() == 0               # pd_create leaves the caller's device alone
        x0 = rand_global(h0.size)
        x1 = x0.to("cuda:1")
        y0 = h0.pc_apply(x0)
        y1 = h1.pc_apply(x1)                                  # current device is 0: the library switches itself
        assert torch.cuda.current_device() == 0
        assert torch.equal(y0.cpu(), y1.cpu())
        b1 = h1.build_rhs(torch.empty(h1.size, dtype=torch.complex128, device="cuda:1"))
        s1, its, _, reason = h1.gmres(b1, x=torch.empty_like(b1), rtol=1e-7)
        assert reason == "CONVERGED_RTOL" and torch.cuda.current_device() == 0
        assert np.array_equal(h1.pc_apply_host(x0.cpu().numpy()), y0.cpu().numpy())
    # one slab per device, plain peer access, one process
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, 2, devices=[0, 1]) as grp:
        for rep in range(3):
            x = rand_global(h.size, seed=rep)
            assert relerr(grp.apply(x), h.pc_apply(x)) < 1e-10
        assert all(not to for to, _ in grp.status())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, N_x, N_t, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from optimal_control_paradiag_b200.dist import DistributedDiagFFTPC
        dpc = DistributedDiagFFTPC(N_x, N_t, device=rank, mode="slab")
        errs = []
        with ParaDiagHandle(N_x, N_t, device=rank) as h:
            for rep in range(6):
                g = torch.Generator(device=f"cuda:{rank}").manual_seed(rep)
                xg = torch.randn(h.size, dtype=torch.float64, device=f"cuda:{rank}", generator=g) + 0j
                xg = xg + 1j * torch.randn(h.size, dtype=torch.float64, device=f"cuda:{rank}", generator=g)
                ref = h.pc_apply(xg)
                if rep % 2 == rank % 2:
                    time.sleep(0.05)                           # skew the ranks: the waits must really wait
                yl = dpc.apply(dpc.scatter_from_global(xg))
                yg = dpc.gather_to_global(yl)
                errs.append(float(torch.linalg.norm(yg - ref) / torch.linalg.norm(ref)))
            xr = torch.randn(h.size, dtype=torch.float64, device=f"cuda:{rank}",
                             generator=torch.Generator(device=f"cuda:{rank}").manual_seed(77))
            if h.real_path_supported:
                refr = h.pc_apply_real(xr).view(2, N_x + 1, N_t)[:, dpc.noff[rank]:dpc.noff[rank + 1], :].reshape(-1)
                yr = dpc.apply_real(xr.view(2, N_x + 1, N_t)[:, dpc.noff[rank]:dpc.noff[rank + 1], :].reshape(-1).contiguous())
                errs.append(float(torch.linalg.norm(yr - refr) / torch.linalg.norm(refr)))
        dpc.check_exchange()
        ret[rank] = (dpc.transport, max(errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N_x,N_t", [(255, 128), (1024, 1024), (4096, 4096), (40, 16384), (80, 81)])
def test_peer_exchange_across_processes_over_nvlink(N_x, N_t):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    ret = mp.Manager().dict()
    mp.spawn(_ipc_worker, args=(world, _free_port(), N_x, N_t, ret), nprocs=world, join=True)
    for r in range(world):
        transport, err = ret[r]
        assert transport == "peer", ret[r]
        assert err < 1e-10, (r, err)


@pytest.mark.parametrize("N_x,N_t", [(255, 256), (1024, 1024), (80, 81)])
def test_single_gpu_apply_replays_from_a_cuda_graph(N_x, N_t):
    # the apply is launch-only once its work buffers exist (no allocation, no host synchronisation), with or without
    # the programmatic-dependent-launch attribute: capture once, replay on new data
    with ParaDiagHandle(N_x, N_t) as h:
        x = rand_global(h.size, seed=1)
        y = torch.empty_like(x)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            h.pc_apply(x, y)                                   # warm-up on a side stream: buffers, attributes
        torch.cuda.current_stream().wait_stream(side)
        ref = h.pc_apply(x)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            h.pc_apply(x, y)
        for seed in (1, 2, 3):
            x.copy_(rand_global(h.size, seed=seed))
            want = h.pc_apply(x)
            y.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(y, want), seed
        assert torch.equal(ref, h.pc_apply(rand_global(h.size, seed=1)))


@pytest.mark.parametrize("N_x,N_t,G,real", [(255, 128, 3, False), (1024, 1024, 4, False), (1024, 256, 4, True)])
def test_slab_apply_with_peer_exchange_replays_from_a_cuda_graph(N_x, N_t, G, real):
    # epochs, parities and flags of the exchange live in device memory: the whole distributed apply (all ranks of a
    # LocalSlabGroup on this GPU) is captured once and replayed; every replay is a new epoch
    with ParaDiagHandle(N_x, N_t) as h, LocalSlabGroup(N_x, N_t, G) as grp:
        x = rand_global(h.size, seed=1, real=real)
        xs = grp.scatter(x)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            grp.apply_blocks(xs, real=real)                    # warm-up: epoch 1
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ys = grp.apply_blocks(xs, real=real)               # capture launches nothing: the epoch stays at 1
        for rep, seed in enumerate((4, 5, 6)):
            xn = rand_global(h.size, seed=seed, real=real)
            for dst, src in zip(xs, grp.scatter(xn)):
                dst.copy_(src)
            g.replay()
            torch.cuda.synchronize()
            y = torch.cat([b.view(2, grp.ncount[r], N_t) for r, b in enumerate(ys)], dim=1).reshape(-1)
            want = h.pc_apply_real(xn) if real else h.pc_apply(xn)
            assert relerr(y, want) < 1e-10, (rep, relerr(y, want))
        assert all(not to and ep == 4 for to, ep in grp.status()), grp.status()
