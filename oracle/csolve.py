"""ctypes binding of oracle/csrc/pc_solve.c (pthread-parallel batched Thomas) -- oracle code.

Used as the ``solver`` of ``DiagFFTPCFast`` for the timed CPU baseline so that the baseline
uses every host core; numerically the same recurrence as ``pc_fast.thomas_toeplitz``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_solve.so")
_lib = None


def build():
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True, capture_output=True)


def load(build_if_missing=True):
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            if not build_if_missing:
                raise FileNotFoundError(_SO)
            build()
        lib = C.CDLL(_SO)
        lib.oracle_thomas_toeplitz.restype = C.c_int
        lib.oracle_thomas_toeplitz.argtypes = [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_pc_stage.restype = C.c_int
        lib.oracle_pc_stage.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


def num_threads():
    return load().oracle_num_threads()


def set_num_threads(n):
    load().oracle_set_num_threads(int(n))


def thomas_toeplitz_c(a, b, rhs, conj_mode=False):
    """In-place capable Thomas solve over all frequencies; rhs (m, K) complex128."""
    lib = load()
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    out = np.array(rhs, dtype=np.complex128, order="C", copy=True)
    m, K = out.shape
    rc = lib.oracle_thomas_toeplitz(m, K, K, a.ctypes.data, b.ctypes.data, out.ctypes.data, int(conj_mode))
    if rc:
        raise MemoryError("oracle_thomas_toeplitz")
    return out


def pc_stage_c(a, b, z, sigma, xh):
    """The whole per-frequency stage (rotation in, both Thomas solves, rotation out, Dirichlet rows), fused and
    threaded, IN PLACE on xh = ifft_t(x) of shape (2, n, K) complex128 (Control_Wave_PC.py:445-540)."""
    lib = load()
    assert xh.dtype == np.complex128 and xh.flags.c_contiguous and xh.ndim == 3 and xh.shape[0] == 2
    a = np.ascontiguousarray(a, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    z = np.ascontiguousarray(z, dtype=np.complex128)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    rc = lib.oracle_pc_stage(xh.shape[1], xh.shape[2], a.ctypes.data, b.ctypes.data, z.ctypes.data,
                             sigma.ctypes.data, xh.ctypes.data)
    if rc:
        raise MemoryError(f"oracle_pc_stage: {rc}")
    return xh
