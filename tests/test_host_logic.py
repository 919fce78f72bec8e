"""Host-side mirror of the reference interface: configuration lookup, the python-PC protocol,
error behaviour.  No GPU: the handle is replaced by a recording fake."""
import numpy as np
import pytest

import optimal_control_paradiag_b200 as pkg
from optimal_control_paradiag_b200 import pc as pcmod
from optimal_control_paradiag_b200 import petsc_shim, problem


class FakeHandle:
    created = []

    def __init__(self, N_x, N_t, T=2.0, gamma=1.0, bug138=True, device=0, **kw):
        self.N_x, self.N_t, self.T, self.gamma, self.device = N_x, N_t, T, gamma, device
        self.n = N_x + 1
        self.size = 2 * self.n * N_t
        self.alpha = kw.get("alpha", 1.0)
        self.calls = []
        FakeHandle.created.append(self)

    def pc_apply_host(self, x, y=None):
        self.calls.append("host")
        out = 2.0 * np.asarray(x).reshape(-1)
        if y is not None:
            y[...] = out
            return y
        return out

    def close(self):
        self.calls.append("close")


@pytest.fixture(autouse=True)
def fake_handle(monkeypatch):
    FakeHandle.created.clear()
    monkeypatch.setattr(pcmod, "ParaDiagHandle", FakeHandle)
    monkeypatch.setattr(pkg.DiagFFTPC, "_defaults", {})
    yield


def test_class_surface_matches_reference():
    for name in ("initialize", "update", "apply", "applyTranspose", "setUp"):
        assert callable(getattr(pkg.DiagFFTPC, name))
    assert pkg.DiagFFTPC.__name__ == "DiagFFTPC"
    with pytest.raises(NotImplementedError):
        pkg.DiagFFTPC().applyTranspose(None, None, None)
    assert pkg.DiagFFTPC().update(None) is None


def test_config_from_configure_and_apply_through_pc():
    pkg.DiagFFTPC.configure(N_x=4, N_t=3, T=2.0, gamma=0.5)
    pc = petsc_shim.PC()
    pc.setPythonType("optimal_control_paradiag_b200.DiagFFTPC")     # like -pc_python_type
    pc.setUp()
    h = FakeHandle.created[-1]
    assert (h.N_x, h.N_t, h.T, h.gamma) == (4, 3, 2.0, 0.5)
    x = petsc_shim.Vec(np.arange(2 * 5 * 3) + 1j)
    y = petsc_shim.Vec.zeros(x.getSize())
    pc.apply(x, y)
    assert np.allclose(y.getArray(), 2 * x.getArray())
    pc.setUp()                      # later set-ups call update(), not initialize()
    assert len(FakeHandle.created) == 1
    pc.destroy()
    assert "close" in h.calls


def test_options_prefix_overrides_configure():
    pkg.DiagFFTPC.configure(N_x=4, N_t=3, T=2.0, gamma=0.5)
    opts = petsc_shim.Options({"firedrake_0_diagfft_nx": 8, "firedrake_0_diagfft_nt": 5,
                               "firedrake_0_diagfft_gamma": "1e-2", "firedrake_0_diagfft_device": 0})
    pc = petsc_shim.PC(prefix="firedrake_0_", options=opts)
    ctx = pkg.DiagFFTPC()
    pc.setPythonContext(ctx)
    pc.setUp()
    h = FakeHandle.created[-1]
    assert (h.N_x, h.N_t, h.gamma, h.T) == (8, 5, 0.01, 2.0)


def test_alpha_defaults_to_the_upstream_operator_and_follows_the_option_chain():
    # the reference has no alpha: 1.0 unless explicitly asked for (configure < options prefix)
    pkg.DiagFFTPC.configure(N_x=4, N_t=3, T=2.0, gamma=0.5)
    pc = petsc_shim.PC()
    pc.setPythonContext(pkg.DiagFFTPC())
    pc.setUp()
    assert FakeHandle.created[-1].alpha == 1.0
    pkg.DiagFFTPC.configure(alpha=0.25)
    pc = petsc_shim.PC()
    pc.setPythonContext(pkg.DiagFFTPC())
    pc.setUp()
    assert FakeHandle.created[-1].alpha == 0.25
    pc = petsc_shim.PC(prefix="p_", options=petsc_shim.Options({"p_diagfft_alpha": "1e-3"}))
    pc.setPythonContext(pkg.DiagFFTPC())
    pc.setUp()
    assert FakeHandle.created[-1].alpha == 1e-3
    with pytest.raises(TypeError):
        pkg.DiagFFTPC.configure(beta=1.0)


def test_appctx_has_highest_priority(monkeypatch):
    pkg.DiagFFTPC.configure(N_x=4, N_t=3, T=2.0, gamma=0.5)
    monkeypatch.setattr(pkg.DiagFFTPC, "get_appctx", staticmethod(lambda pc: {"paradiag": {"N_x": 6, "N_t": 7}}))
    ctx = pkg.DiagFFTPC()
    ctx.setUp(petsc_shim.PC())
    assert (FakeHandle.created[-1].N_x, FakeHandle.created[-1].N_t) == (6, 7)


def test_globals_of_main_like_the_reference_script(monkeypatch):
    import sys
    main = sys.modules["__main__"]
    for k, v in dict(N_x=10, N_t=9, T=2, gamma=1).items():
        monkeypatch.setattr(main, k, v, raising=False)
    ctx = pkg.DiagFFTPC()
    ctx.setUp(petsc_shim.PC())
    assert (FakeHandle.created[-1].N_x, FakeHandle.created[-1].N_t) == (10, 9)


def test_incomplete_config_raises():
    with pytest.raises(ValueError, match="incomplete"):
        pkg.DiagFFTPC().setUp(petsc_shim.PC())


def test_wrong_vec_size_raises():
    pkg.DiagFFTPC.configure(N_x=4, N_t=3, T=2.0, gamma=1.0)
    ctx = pkg.DiagFFTPC()
    pc = petsc_shim.PC()
    ctx.setUp(pc)
    with pytest.raises(ValueError, match="Vec size"):
        ctx.apply(pc, petsc_shim.Vec(np.zeros(7)), petsc_shim.Vec(np.zeros(7)))


def test_node_order_permutation():
    N_x, N_t = 4, 3
    order = np.array([1, 0, 2, 4, 3])         # Vec position -> geometric node
    pkg.DiagFFTPC.configure(N_x=N_x, N_t=N_t, T=2.0, gamma=1.0, node_order=order)
    ctx = pkg.DiagFFTPC()
    pc = petsc_shim.PC()
    ctx.setUp(pc)
    seen = {}
    h = FakeHandle.created[-1]
    orig = h.pc_apply_host

    def spy(x, y=None):
        seen["x"] = np.array(x).reshape(2, N_x + 1, N_t).copy()
        return orig(x, y)
    h.pc_apply_host = spy
    x = np.arange(2 * (N_x + 1) * N_t, dtype=complex)
    y = petsc_shim.Vec.zeros(x.size)
    ctx.apply(pc, petsc_shim.Vec(x), y)
    xs = x.reshape(2, N_x + 1, N_t)
    for pos, geo in enumerate(order):
        assert np.array_equal(seen["x"][:, geo, :], xs[:, pos, :])
    assert np.allclose(y.getArray(), 2 * x)   # un-permuted on the way out
    with pytest.raises(ValueError, match="permutation"):
        pkg.DiagFFTPC.configure(node_order=[0, 0, 1, 2, 3])
        pkg.DiagFFTPC().setUp(petsc_shim.PC())


def test_reference_solver_parameters_are_the_default():
    p = problem.default_parameters
    assert p["ksp_type"] == "gmres" and p["ksp_gmres_restart"] == 300 and p["ksp_max_it"] == 1000
    assert p["pc_type"] == "python" and p["pc_python_type"].endswith(".DiagFFTPC")
    assert p["snes_type"] == "ksponly" and p["mat_type"] == "matfree"
    flat = problem._flatten(p)
    assert "ksp_monitor" in flat and "ksp_converged_reason" in flat


def test_petsc_shim_vec_and_options():
    v = petsc_shim.Vec(np.arange(4))
    ro = v.getArray(readonly=True)
    with pytest.raises(ValueError):
        ro[0] = 1
    w = v.duplicate()
    v.copy(w)
    assert np.array_equal(w.getArray(), v.getArray()) and v.getSize() == 4
    o = petsc_shim.Options({"-a_b": 3})
    assert o.hasName("a_b") and o.getInt("a_b") == 3 and o.getReal("zz", 1.5) == 1.5
    pc = petsc_shim.PC(options=petsc_shim.Options({"pc_type": "python",
                                                   "pc_python_type": "optimal_control_paradiag_b200.DiagFFTPC"}))
    pc.setFromOptions()
    assert isinstance(pc.getPythonContext(), pkg.DiagFFTPC)
