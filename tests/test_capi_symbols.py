"""The C-ABI library loads on a CPU-only box and exports every symbol include/paradiag.h declares
(no compute calls without a GPU).  It must also fail loudly -- not fall back -- without a device."""
import ctypes as C
import os
import re

import pytest

import optimal_control_paradiag_b200 as pkg
from optimal_control_paradiag_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "paradiag.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    path = pkg.library_path()
    assert os.path.exists(path), f"{path} missing: run __graft_entry__.build()"
    assert os.path.dirname(path) == os.path.join(ROOT, "optimal_control_paradiag_b200")


def test_every_declared_symbol_is_exported_and_bound():
    lib = pkg.load_library()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"libparadiag.so does not export {s}"
        assert s in _lib.SYMBOLS, f"{s} is declared in paradiag.h but not bound in _lib.py"
    for s in _lib.SYMBOLS:
        assert s in syms, f"{s} is bound in _lib.py but not declared in paradiag.h"


def test_abi_version_and_struct_size():
    lib = pkg.load_library()
    assert lib.pd_abi_version() == _lib.PD_ABI_VERSION
    # pd_config: 4 int32, 3 double, 4 int32, 5 reserved int32 -> 16 + 24 + 16 + 20 = 76 -> padded to 80
    assert C.sizeof(_lib.pd_config) == 80


def test_invalid_config_is_rejected_before_touching_cuda():
    lib = pkg.load_library()
    h = C.c_void_p()
    cfg = _lib.pd_config(abi_version=999, N_x=8, N_t=8, T=2.0, gamma=1.0, alpha=1.0)
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == _lib.PD_ERR_INVALID
    assert b"ABI" in lib.pd_last_error()
    cfg = _lib.pd_config(abi_version=_lib.PD_ABI_VERSION, N_x=1, N_t=8, T=2.0, gamma=1.0, alpha=1.0)
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == _lib.PD_ERR_INVALID
    # alpha outside (0, 1] is invalid; alpha != 1 (an extension) on a frequency-sharded stage handle is unsupported
    cfg = _lib.pd_config(abi_version=_lib.PD_ABI_VERSION, N_x=8, N_t=8, T=2.0, gamma=1.0, alpha=1.5)
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == _lib.PD_ERR_INVALID
    cfg = _lib.pd_config(abi_version=_lib.PD_ABI_VERSION, N_x=8, N_t=8, T=2.0, gamma=1.0, alpha=0.0)
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == _lib.PD_ERR_INVALID
    cfg = _lib.pd_config(abi_version=_lib.PD_ABI_VERSION, N_x=8, N_t=8, T=2.0, gamma=1.0, alpha=0.1,
                         k_begin=0, k_count=4)
    assert lib.pd_create(C.byref(cfg), C.byref(h)) == _lib.PD_ERR_UNSUPPORTED
    assert lib.pd_create(None, C.byref(h)) == _lib.PD_ERR_INVALID
    assert lib.pd_destroy(None) == 0
    assert lib.pd_pc_apply_transpose(None, None, None, None) == _lib.PD_ERR_UNSUPPORTED


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.ParaDiagError) as ei:
        pkg.ParaDiagHandle(16, 16)
    assert ei.value.status == _lib.PD_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("PARADIAG_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(pkg.LibraryNotBuilt):
        pkg.load_library()
    monkeypatch.delenv("PARADIAG_LIB")
    monkeypatch.setattr(_lib, "_lib", None)
    pkg.load_library()


def test_product_package_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "optimal_control_paradiag_b200")
    for f in os.listdir(os.path.join(ROOT, "tools")):            # dev tools outside tests/ stay oracle-free too
        if f.endswith((".py", ".sh")):
            assert not re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, "tools", f)).read(),
                                 flags=re.M), f
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def _header_config_fields():
    """(name, C type, array length) of every member of `struct pd_config` in include/paradiag.h."""
    import re
    src = open(os.path.join(ROOT, "include", "paradiag.h")).read()
    body = src[src.index("typedef struct pd_config {"):src.index("} pd_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for m in re.finditer(r"\b(int32_t|double)\s+(\w+)(?:\[(\d+)\])?\s*;", body):
        out.append((m.group(2), m.group(1), int(m.group(3)) if m.group(3) else 0))
    return out


def test_pd_config_mirrors_match_the_header_field_for_field():
    # the ctypes mirror of the package AND the inline stub of INTEGRATION.md section 2 against include/paradiag.h
    import ctypes as C
    import re
    fields = _header_config_fields()
    assert [f[0] for f in fields][:4] == ["abi_version", "N_x", "N_t", "bug138"] and fields[-1][0] == "reserved"
    ctype = {"int32_t": C.c_int32, "double": C.c_double}
    want = [(n, ctype[t] * ln if ln else ctype[t]) for n, t, ln in fields]
    got = list(_lib.pd_config._fields_)
    assert [g[0] for g in got] == [w[0] for w in want]
    for (gn, gt), (wn, wt) in zip(got, want):
        assert C.sizeof(gt) == C.sizeof(wt), gn
    # the names inside its `_fields_ = [...]` list, in order
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = doc[doc.index("class pd_config(C.Structure):"):doc.index("lib = C.CDLL(")]
    names = re.findall(r'\("(\w+)",\s*C\.c_(?:int32|double)(?:\s*\*\s*(\d+))?\)', stub)
    assert [n for n, _ in names] == [f[0] for f in fields]
    assert [int(l) if l else 0 for _, l in names] == [f[2] for f in fields]


def test_hessenberg_state_matches_a_dense_least_squares_solve():
    # pd_hess_* (host only): Givens-reduced Hessenberg of a GMRES cycle against numpy's least squares on the same H
    import numpy as np
    from optimal_control_paradiag_b200._lib import Hessenberg
    rng = np.random.default_rng(0)
    m, beta = 9, 2.5
    H = np.zeros((m + 1, m), dtype=complex)
    q = Hessenberg(m)
    q.start(beta)
    for j in range(m):
        col = rng.standard_normal(j + 1) + 1j * rng.standard_normal(j + 1)
        hn = abs(rng.standard_normal()) + 0.1
        H[: j + 1, j], H[j + 1, j] = col, hn
        rn, hn_out = q.push(np.concatenate([col, [hn * hn]]))
        assert abs(hn_out - hn) < 1e-15
        e1 = np.zeros(j + 2, dtype=complex)
        e1[0] = beta
        y, *_ = np.linalg.lstsq(H[: j + 2, : j + 1], e1, rcond=None)
        assert abs(rn - np.linalg.norm(e1 - H[: j + 2, : j + 1] @ y)) < 1e-12      # residual estimate = LS residual
    assert np.allclose(q.solve(), y, rtol=1e-11, atol=1e-13)
    q.start(1.0)                                                                  # a new cycle reuses the state
    rn, _ = q.push(np.array([3.0 + 0j, 16.0]))
    assert abs(rn - 4.0 / 5.0) < 1e-15 and np.allclose(q.solve(), [3.0 / 25.0])
    lib = pkg.load_library()
    assert lib.pd_hess_create(0, None) == _lib.PD_ERR_INVALID
    q.close()
