"""ctypes binding of ``libparadiag.so`` (the C ABI of ``include/paradiag.h``).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make`` in ``csrc/``).
Loading fails loudly when it is missing: the product path has no fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = "libparadiag.so"
PD_ABI_VERSION = 1

PD_OK, PD_ERR_INVALID, PD_ERR_CUDA, PD_ERR_NOMEM, PD_ERR_UNSUPPORTED, PD_ERR_NOT_CONVERGED = 0, -1, -2, -3, -4, -5


class LibraryNotBuilt(RuntimeError):
    pass


class ParaDiagError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libparadiag error {status}: {message}")
        self.status = status


class pd_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("N_x", C.c_int32), ("N_t", C.c_int32), ("bug138", C.c_int32),
        ("T", C.c_double), ("gamma", C.c_double), ("alpha", C.c_double),
        ("device", C.c_int32), ("k_begin", C.c_int32), ("k_count", C.c_int32), ("n_local", C.c_int32),
        ("slab_rank", C.c_int32), ("slab_count", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


# every symbol include/paradiag.h declares: name -> (restype, argtypes)
_VP, _I, _I64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = {
    "pd_create": (_I, [C.POINTER(pd_config), C.POINTER(_VP)]),
    "pd_destroy": (_I, [_VP]),
    "pd_last_error": (C.c_char_p, []),
    "pd_abi_version": (_I, []),
    "pd_workspace_bytes": (C.c_size_t, [_VP]),
    "pd_launch_count": (_I64, [_VP]),
    "pd_pc_apply": (_I, [_VP, _VP, _VP, _VP]),
    "pd_pc_apply_host": (_I, [_VP, _VP, _VP]),
    "pd_host_unregister_all": (_I, [_VP]),
    "pd_pc_apply_real": (_I, [_VP, _VP, _VP, _VP]),
    "pd_stage_rfft": (_I, [_VP, _VP, _VP, _I64, _I, _VP]),
    "pd_stage_solve_half": (_I, [_VP, _VP, _VP]),
    "pd_pc_apply_profile": (_I, [_VP, _VP, _VP, _VP, C.POINTER(C.c_float), _I]),
    "pd_pc_apply_transpose": (_I, [_VP, _VP, _VP, _VP]),
    "pd_stage_fft": (_I, [_VP, _VP, _VP, _I64, _I, _VP]),
    "pd_stage_gamma": (_I, [_VP, _VP, _VP, _I64, _I, _VP]),
    "pd_stage_solve": (_I, [_VP, _VP, _VP]),
    "pd_slab_reduce": (_I, [_VP, _VP, _VP, _VP]),
    "pd_slab_finish": (_I, [_VP, _VP, _VP, _VP]),
    "pd_stage_rfft_pair": (_I, [_VP, _VP, _VP, _I64, _I, _VP]),
    "pd_slab_reduce_half": (_I, [_VP, _VP, _VP, _VP]),
    "pd_slab_finish_half": (_I, [_VP, _VP, _VP, _VP]),
    "pd_slab_comm_create": (_I, [_VP, _VP, C.POINTER(_VP)]),
    "pd_slab_comm_connect": (_I, [_VP, _VP, _I, C.POINTER(_I)]),
    "pd_slab_comm_status": (_I, [_VP, C.POINTER(_I), C.POINTER(C.c_uint64)]),
    "pd_slab_apply": (_I, [_VP, _VP, _VP, _VP]),
    "pd_slab_apply_real": (_I, [_VP, _VP, _VP, _VP]),
    "pd_slab_apply_begin": (_I, [_VP, _VP, _VP, _I]),
    "pd_slab_apply_end": (_I, [_VP, _VP, _VP, _I]),
    "pd_slab_apply_profile": (_I, [_VP, _VP, _VP, _VP, C.POINTER(C.c_float), _I]),
    "pd_matvec": (_I, [_VP, _VP, _VP, _VP]),
    "pd_matvec_slab": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "pd_matvec_slab_real": (_I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "pd_pc_apply_real_host": (_I, [_VP, _VP, _VP]),
    "pd_pc_matvec": (_I, [_VP, _VP, _VP, _VP]),
    "pd_build_rhs": (_I, [_VP, _VP, _VP]),
    "pd_gmres": (_I, [_VP, _VP, _VP, _D, _D, _I, _I, C.POINTER(_I), C.POINTER(_D), C.POINTER(_I), _VP]),
    "pd_set_option": (_I, [_VP, C.c_char_p, _D]),
    "pd_delta": (_I, [_VP, _VP, _VP, _I, _VP]),
    "pd_matvec_real": (_I, [_VP, _VP, _VP, _VP]),
    "pd_build_rhs_real": (_I, [_VP, _VP, _VP]),
    "pd_gmres_real": (_I, [_VP, _VP, _VP, _D, _D, _I, _I, C.POINTER(_I), C.POINTER(_D), C.POINTER(_I), _VP]),
    "pd_mdot": (_I, [_VP, _VP, _I64, _I, _VP, _I64, _VP, _VP]),
    "pd_maxpy": (_I, [_VP, _VP, _I64, _I, _VP, _D, _VP, _I64, _VP, _VP]),
    "pd_hess_create": (_I, [_I, C.POINTER(_VP)]),
    "pd_hess_destroy": (_I, [_VP]),
    "pd_hess_start": (_I, [_VP, _D]),
    "pd_hess_push": (_I, [_VP, _VP, C.POINTER(_D), C.POINTER(_D)]),
    "pd_hess_solve": (_I, [_VP, _VP, C.POINTER(_I)]),
}

_lib = None


def library_path():
    return os.environ.get("PARADIAG_LIB", os.path.join(_HERE, _LIBNAME))


def load_library():
    """Load libparadiag.so and type every entry point.  Raises LibraryNotBuilt if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LibraryNotBuilt(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C optimal_control_paradiag_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pd_abi_version() != PD_ABI_VERSION:
        raise LibraryNotBuilt(f"{path}: ABI version {lib.pd_abi_version()} != {PD_ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status, allow=()):
    if status != PD_OK and status not in allow:
        raise ParaDiagError(status, load_library().pd_last_error().decode())
    return status


class Hessenberg:
    """Host-side state of one restarted-GMRES cycle (``pd_hess_*`` in include/paradiag.h): the same Givens /
    Hessenberg recurrence ``pd_gmres`` runs internally, for Krylov loops that own their collectives (dist.py)."""

    def __init__(self, restart):
        import numpy as np
        self.np = np
        self.lib = load_library()
        self.restart = int(restart)
        self._q = C.c_void_p()
        check(self.lib.pd_hess_create(self.restart, C.byref(self._q)))

    def start(self, beta):
        check(self.lib.pd_hess_start(self._q, float(beta)))

    def push(self, hcol):
        """hcol: complex array (h_0..h_j, squared norm of the orthogonalised vector); returns (|g_{j+1}|, h_{j+1,j})."""
        a = self.np.ascontiguousarray(hcol, dtype=self.np.complex128)
        rn, hn = _D(), _D()
        check(self.lib.pd_hess_push(self._q, a.ctypes.data_as(_VP), C.byref(rn), C.byref(hn)))
        return rn.value, hn.value

    def solve(self):
        y = self.np.empty(self.restart, dtype=self.np.complex128)
        n = _I()
        check(self.lib.pd_hess_solve(self._q, y.ctypes.data_as(_VP), C.byref(n)))
        return y[:n.value].copy()

    def close(self):
        if self._q:
            self.lib.pd_hess_destroy(self._q)
            self._q = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
