"""World-size-2/3 gloo tests (CPU) of the multi-GPU host logic: slab partition, pack/unpack and
the two all-to-all transposes of ``DistributedDiagFFTPC``.  The compute stages are stood in by the
oracle (tests only); on the GPU box the same class drives libparadiag (tests/test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from optimal_control_paradiag_b200.dist import DistributedDiagFFTPC, slab_bounds  # noqa: E402


class OracleStageBackend:
    """pd_stage_fft / pd_stage_solve semantics on CPU tensors, from the oracle's fast route."""
    launch_count = 0

    def __init__(self, N_x, N_t, T, gamma, k_begin, k_count, n_local):
        from oracle.pc_fast import DiagFFTPCFast
        self.pc = DiagFFTPCFast(N_x, N_t, T, gamma, workers=1)
        self.N_t, self.n = N_t, N_x + 1
        self.ks = slice(k_begin, k_begin + k_count)
        self.k_count = k_count

    def stage_fft(self, src, dst, nlines, inverse):
        import scipy.fft as sfft
        a = src.numpy().reshape(nlines, self.N_t)
        out = sfft.ifft(a, axis=1) if inverse else sfft.fft(a, axis=1)
        dst.copy_(torch.from_numpy(out.reshape(-1)))

    def stage_solve(self, w):
        pc, ks = self.pc, self.ks
        W = w.numpy().reshape(2, self.n, self.k_count)
        z, sg = pc.z[ks], pc.sigma[ks]
        uz = W[0] * np.conj(z)
        ip = (1j * sg) * W[1]
        rp, rm = (uz + ip) / 2, (uz - ip) / 2
        zp, zm = np.zeros_like(rp), np.zeros_like(rm)
        zp[1:-1] = pc.solver(pc.a[ks], pc.b[ks], rp[1:-1])
        zm[1:-1] = np.conj(pc.solver(pc.a[ks], pc.b[ks], np.conj(rm[1:-1])))
        W[0] = zp + zm
        W[1] = (-1j * sg * z) * (zp - zm)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N_x, N_t, gamma, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pc_fast import DiagFFTPCFast
        factory = lambda **kw: OracleStageBackend(N_x, N_t, 2.0, gamma, **kw)
        dpc = DistributedDiagFFTPC(N_x, N_t, T=2.0, gamma=gamma, backend_factory=factory)
        rng = np.random.default_rng(0)
        size = 2 * (N_x + 1) * N_t
        xg = rng.standard_normal(size) + 1j * rng.standard_normal(size)
        x_local = dpc.scatter_from_global(torch.from_numpy(xg))
        y_local = dpc.apply(x_local)
        yg = dpc.gather_to_global(y_local).numpy()
        ref = DiagFFTPCFast(N_x, N_t, 2.0, gamma).apply(xg)
        err = np.linalg.norm(yg - ref) / np.linalg.norm(ref)
        d = dpc.describe()
        ok = err < 1e-12 and sum(d["node_slabs"]) == N_x + 1 and sum(d["freq_slabs"]) == N_t
        ret[rank] = (bool(ok), float(err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N_x,N_t", [(2, 16, 12), (2, 21, 9), (3, 20, 16)])
def test_distributed_apply_matches_single_process_oracle(world, N_x, N_t):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), N_x, N_t, 0.5, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, err = ret[r]
        assert ok, (r, err)


def test_slab_bounds():
    assert slab_bounds(16385, 8) == ([2049] + [2048] * 7, [0, 2049, 4097, 6145, 8193, 10241, 12289, 14337, 16385])
    c, o = slab_bounds(10, 3)
    assert c == [4, 3, 3] and o == [0, 4, 7, 10]
    c, o = slab_bounds(4096, 8)
    assert c == [512] * 8 and o[-1] == 4096
