"""Minimal stand-ins for the petsc4py objects the python-PC protocol touches.

petsc4py / Firedrake are not installable in this image, so the ``DiagFFTPC`` surface is
exercised through these: a ``Vec`` with ``getArray / setArray / copy / getSize / duplicate``,
an ``Options`` database and a ``PC`` with ``getOptionsPrefix``, ``setType('python')``,
``setPythonContext`` / ``setPythonType`` (dotted-path import like PETSc's ``pc_python_type``,
Control_Wave_PC.py:357-358), ``setUp`` and ``apply``.
"""
import importlib

import numpy as np


class Vec:
    array_is_view = True

    def __init__(self, array):
        self._a = np.ascontiguousarray(array, dtype=np.complex128).reshape(-1)

    @classmethod
    def zeros(cls, size):
        return cls(np.zeros(size, dtype=np.complex128))

    def getArray(self, readonly=False):
        if readonly:
            v = self._a.view()
            v.flags.writeable = False
            return v
        return self._a

    def setArray(self, a):
        self._a[...] = np.asarray(a).reshape(-1)

    def getSize(self):
        return self._a.size

    def duplicate(self):
        return Vec(np.zeros_like(self._a))

    def copy(self, other=None):
        if other is None:
            return Vec(self._a.copy())
        other.setArray(self._a)
        return other

    def norm(self):
        return float(np.linalg.norm(self._a))


class Comm:
    """Stand-in for a petsc4py communicator: ``getSize`` / ``getRank``."""

    def __init__(self, size=1, rank=0):
        self._size, self._rank = int(size), int(rank)

    def getSize(self):
        return self._size

    def getRank(self):
        return self._rank


class Options:
    def __init__(self, entries=None):
        self._d = {}
        for k, v in (entries or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        self._d[k.lstrip("-")] = None if v is None else str(v)

    def __getitem__(self, k):
        return self._d[k.lstrip("-")]

    def hasName(self, k):
        return k.lstrip("-") in self._d

    def getString(self, k, default=None):
        return self._d.get(k.lstrip("-"), default)

    def getInt(self, k, default=None):
        return int(self._d[k.lstrip("-")]) if self.hasName(k) else default

    def getReal(self, k, default=None):
        return float(self._d[k.lstrip("-")]) if self.hasName(k) else default


class PC:
    """Just enough of ``PETSc.PC`` to drive a python-type preconditioner."""

    def __init__(self, prefix="", options=None, comm=None):
        self._prefix = prefix
        self._comm = comm if comm is not None else Comm()
        self.options = options if options is not None else Options()
        self._ctx = None
        self._type = None
        self._setup = False

    def getOptionsPrefix(self):
        return self._prefix

    def getComm(self):
        return self._comm

    def setOptionsPrefix(self, p):
        self._prefix = p

    def setType(self, t):
        self._type = t

    def getType(self):
        return self._type

    def setPythonContext(self, ctx):
        self._ctx = ctx
        self._type = "python"

    def getPythonContext(self):
        return self._ctx

    def setPythonType(self, dotted):
        """PETSc semantics of ``-pc_python_type module.Class``: import and instantiate."""
        mod, _, cls = dotted.rpartition(".")
        self.setPythonContext(getattr(importlib.import_module(mod), cls)())

    def setFromOptions(self):
        t = self.options.getString(self._prefix + "pc_type")
        if t:
            self._type = t
        pt = self.options.getString(self._prefix + "pc_python_type")
        if pt and self._type == "python":
            self.setPythonType(pt)

    def getDM(self):
        return None

    def setUp(self):
        if self._ctx is None:
            raise RuntimeError("PC of type python has no context")
        self._ctx.setUp(self)
        self._setup = True

    def apply(self, x, y):
        if not self._setup:
            self.setUp()
        self._ctx.apply(self, x, y)

    def applyTranspose(self, x, y):
        if not self._setup:
            self.setUp()
        self._ctx.applyTranspose(self, x, y)

    def destroy(self):
        if self._ctx is not None and hasattr(self._ctx, "destroy"):
            self._ctx.destroy(self)
