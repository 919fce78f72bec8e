"""A stand-in for the narrow slice of the Firedrake / UFL API that ``Code/Control_Wave_PC.py`` uses.

TEST INFRASTRUCTURE (fixture generation only).  Firedrake, PETSc and MUMPS are not installable in this image, so
the upstream script cannot run as it is.  Its classes are nevertheless plain Python that only *talks* to Firedrake:
with this module bound to the name ``fd``, ``make_reference_executed_golden.py`` EXECUTES -- unmodified, read from the
upstream checkout at generation time --

* ``Optimal_Control_Wave_Equation.__init__`` / ``Build_f`` / ``Build_g`` / ``Build_Initial_Condition`` /
  ``Build_L`` (:13-179): the UFL forms of the all-at-once residual, and
* ``DiagFFTPC.initialize`` / ``apply`` (:380-553): the eigen set-up, the forms of the per-frequency block
  systems, the FFTs, the Riesz round trip, the S / S^-1 rotations, the 1/lambda_2 scaling and every copy,

and records what they compute.  What this module supplies instead of Firedrake is deliberately the textbook part:

* P1 elements on the uniform ``UnitIntervalMesh(N_x)``: nodal values, mass matrix h/6 tridiag(1, 4, 1) (2h/6 at the
  two end nodes), stiffness matrix 1/h tridiag(-1, 2, -1) (1/h at the end nodes); ``interpolate`` = nodal evaluation;
* UFL expressions that are AFFINE in coefficient functions (all this script builds): sums, scalar multiples, ``grad``,
  ``inner(., test) * dx`` -> mass- or stiffness-weighted block entries; forms keep references to their coefficient
  ``Function`` objects, whose data is read when a solver runs (as in Firedrake);
* homogeneous ``DirichletBC`` on both fields: boundary rows and columns replaced by the identity, right-hand side 0;
* ``LinearVariationalSolver.solve`` = sparse LU (SuperLU) of the assembled matrix (upstream: MUMPS);
* the mixed-space vector layout [u-block ; p-block], each block (node, component) row-major -- Firedrake's layout for
  ``VectorFunctionSpace(..., dim=N_t) * VectorFunctionSpace(...)`` in serial;
* natural (left-to-right) node numbering.

Nothing here knows about the preconditioner, the time stencils, eigen-decompositions or FFT conventions: those all
come from the executed upstream lines.  Anything the upstream code does not use raises ``NotImplementedError``.
"""
import numbers

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pi = np.pi


# --------------------------------------------------------------------------------------------------- mesh / spaces
class Mesh:
    def __init__(self, N_x):
        self.N_x, self.n = int(N_x), int(N_x) + 1
        self.h = 1.0 / N_x
        self.coords = np.arange(self.n) / float(N_x)
        h, n = self.h, self.n
        main = np.full(n, 4.0)
        main[0] = main[-1] = 2.0
        self.M = sp.diags([np.ones(n - 1), main, np.ones(n - 1)], [-1, 0, 1], format="csr") * (h / 6.0)
        kmain = np.full(n, 2.0)
        kmain[0] = kmain[-1] = 1.0
        self.K = sp.diags([-np.ones(n - 1), kmain, -np.ones(n - 1)], [-1, 0, 1], format="csr") * (1.0 / h)
        self._Mlu = None

    def mass_solve(self, rhs):
        if self._Mlu is None:
            self._Mlu = spla.splu(self.M.tocsc().astype(complex))
        return self._Mlu.solve(np.asarray(rhs, dtype=complex))


def UnitIntervalMesh(N_x):
    return Mesh(N_x)


def UnitSquareMesh(*a, **k):
    raise NotImplementedError("stand-in: 1-D only")


class Space:
    def __init__(self, kind, mesh, dim=None, subs=None):
        self.kind, self.mesh, self.dim, self.subs = kind, mesh, dim, subs

    def __mul__(self, other):
        return Space("MIXED", self.mesh, subs=[self, other])

    def dual(self):
        return self

    def sub(self, i):
        return SubSpace(self, i)


class SubSpace:
    def __init__(self, parent, index):
        self.parent, self.index = parent, index


def FunctionSpace(mesh, family, degree):
    if family == "R":
        return Space("R", mesh)
    if family == "CG" and degree == 1:
        return Space("CG1", mesh)
    raise NotImplementedError(family)


def VectorFunctionSpace(mesh, family, degree, dim):
    assert family == "CG" and degree == 1
    return Space("VEC", mesh, dim=int(dim))


def SpatialCoordinate(mesh):
    return (Nodal(mesh.coords.copy()),)


# ------------------------------------------------------------------------------------------------------- scalars
def _num(v):
    if isinstance(v, Scal):
        return v.value
    if isinstance(v, numbers.Number):
        return v
    return None


class Scal:
    """A spatially constant value: ``Constant``, a ``Function`` on the 'R' space, or arithmetic on those."""

    def __init__(self, value=0.0):
        self.value = value

    def assign(self, v):
        self.value = _num(v)
        return self

    def _bin(self, other, op):
        o = _num(other)
        return NotImplemented if o is None else Scal(op(self.value, o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return Scal(-self.value)


def Constant(v):
    return Scal(v.value if isinstance(v, Scal) else v)


class Nodal:
    """An expression of the spatial coordinate, held by its nodal values (``interpolate`` evaluates at the nodes)."""

    def __init__(self, arr):
        self.arr = np.asarray(arr)

    def _bin(self, other, op):
        if isinstance(other, Nodal):
            return Nodal(op(self.arr, other.arr))
        o = _num(other)
        return NotImplemented if o is None else Nodal(op(self.arr, o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return Nodal(-self.arr)


def _fun(f):
    def g(x):
        if isinstance(x, Nodal):
            return Nodal(f(x.arr))
        v = _num(x)
        if v is None:
            raise NotImplementedError(f"stand-in: {f.__name__} of {type(x).__name__}")
        return Scal(f(v))
    return g


sin, cos, exp, sqrt = _fun(np.sin), _fun(np.cos), _fun(np.exp), _fun(np.sqrt)


# --------------------------------------------------------------------------- expressions affine in functions
class Lin:
    """sum_k coef_k * (component k of a coefficient Function | of the trial function) + offset; ``grad`` marks the
    spatial derivative of the whole expression.  Keys: (function object, component or None) | ('trial', field, comp)."""

    def __init__(self, terms=None, offset=0.0, grad=False):
        self.terms, self.offset, self.grad = dict(terms or {}), offset, grad

    def _lin(self):
        return self

    @staticmethod
    def of(x):
        if isinstance(x, Lin):
            return x
        if hasattr(x, "_lin"):
            return x._lin()
        v = _num(x)
        if v is None:
            raise NotImplementedError(f"stand-in: cannot use {type(x).__name__} in an affine expression")
        return Lin(offset=v)

    def _add(self, other, sign):
        try:
            o = Lin.of(other)
        except NotImplementedError:
            return NotImplemented
        if (self.terms and o.terms and self.grad != o.grad) or (self.grad and o.offset) or (o.grad and self.offset):
            raise NotImplementedError("stand-in: mixing gradients and values in one sum")
        t = dict(self.terms)
        for k, c in o.terms.items():
            t[k] = t.get(k, 0.0) + sign * c
        return Lin(t, self.offset + sign * o.offset, self.grad or o.grad)

    def _scale(self, other, inverse=False):
        v = _num(other)
        if v is None:
            return NotImplemented
        f = 1.0 / v if inverse else v
        return Lin({k: c * f for k, c in self.terms.items()}, self.offset * f, self.grad)

    def __add__(self, o): return self._add(o, 1.0)
    def __radd__(self, o): return self._add(o, 1.0)
    def __sub__(self, o): return self._add(o, -1.0)
    def __rsub__(self, o): return (-self)._add(o, 1.0)
    def __mul__(self, o): return self._scale(o)
    def __rmul__(self, o): return self._scale(o)
    def __truediv__(self, o): return self._scale(o, inverse=True)
    def __neg__(self): return self._scale(-1.0)

    def evaluate(self, n):
        """Nodal values with the coefficient functions' CURRENT data (no trial terms, no gradient)."""
        assert not self.grad
        out = np.full(n, self.offset, dtype=complex)
        for (obj, comp), c in self.terms.items():
            if isinstance(obj, str):
                raise NotImplementedError("stand-in: trial function in a value expression")
            out = out + c * obj.component(comp)
        return out


class LinLike:
    """Mixin: objects that behave as a one-term ``Lin`` in arithmetic."""

    def __add__(self, o): return self._lin() + o
    def __radd__(self, o): return o + self._lin()
    def __sub__(self, o): return self._lin() - o
    def __rsub__(self, o): return o - self._lin()
    def __mul__(self, o): return self._lin() * o
    def __rmul__(self, o): return self._lin() * o
    def __truediv__(self, o): return self._lin() / o
    def __neg__(self): return -self._lin()


class Component(LinLike):
    def __init__(self, key):
        self.key = key

    def _lin(self):
        return Lin({self.key: 1.0})


class VecExpr:
    """``as_vector([...])``: a list of per-component expressions."""

    def __init__(self, items):
        self.items = list(items)

    def __getitem__(self, i):
        return self.items[i]

    def __mul__(self, o):
        return VecExpr([it * o for it in self.items])

    __rmul__ = __mul__


def as_vector(items):
    return VecExpr(items)


# ------------------------------------------------------------------------------------------------------ functions
class Dat:
    def __init__(self, owner):
        self._owner = owner

    @property
    def data(self):
        return self._owner.data

    @property
    def data_ro(self):
        return self._owner.data

    def __getitem__(self, i):
        return Dat(self._owner.subs[i])

    # the mixed vector [u-block ; p-block], each (node, component) row-major
    class _View:
        def __init__(self, fn):
            self.fn = fn

        def setArray(self, arr):
            arr = np.asarray(arr).reshape(-1)
            o = 0
            for s in self.fn.subs:
                s.data[...] = arr[o:o + s.data.size].reshape(s.data.shape)
                o += s.data.size

        def getArray(self, readonly=False):
            return np.concatenate([s.data.reshape(-1) for s in self.fn.subs])

        def copy(self, dest):                       # PETSc Vec.copy(dest): dest <- self
            dest.setArray(self.getArray())

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    @property
    def vec_wo(self):
        return Dat._View(self._owner)

    vec_ro = vec_wo
    vec = vec_wo


class HostVec:
    """The PETSc ``Vec`` the KSP hands to ``apply(pc, x, y)``: an array with ``copy(dest)`` / ``setArray``."""

    def __init__(self, arr):
        self.array = np.array(arr, dtype=complex).reshape(-1)

    def copy(self, dest):
        dest.setArray(self.array)

    def setArray(self, arr):
        self.array[...] = np.asarray(arr).reshape(-1)

    def getArray(self, readonly=False):
        return self.array


class Function(LinLike):
    def __new__(cls, space, name=None):
        if space.kind == "R":
            return Scal(0.0)
        return super().__new__(cls)

    def __init__(self, space, name=None):
        self.space, self.mesh = space, space.mesh
        n = self.mesh.n
        if space.kind == "CG1":
            self.data, self.subs = np.zeros(n, dtype=complex), None
        elif space.kind == "VEC":
            self.data, self.subs = np.zeros((n, space.dim), dtype=complex), None
        elif space.kind == "MIXED":
            self.subs, self.data = [Function(s) for s in space.subs], None
        else:
            raise NotImplementedError(space.kind)

    # --- as an expression
    def _lin(self):
        if self.space.kind != "CG1":
            raise NotImplementedError("stand-in: a vector function in scalar arithmetic")
        return Lin({(self, None): 1.0})

    def __getitem__(self, i):
        assert self.space.kind == "VEC"
        return Component((self, int(i)))

    def component(self, comp):
        return self.data if comp is None else self.data[:, comp]

    # --- data
    @property
    def dat(self):
        return Dat(self)

    @property
    def subfunctions(self):
        return tuple(self.subs)

    def sub(self, i):
        return self.subs[i]

    def assign(self, other):
        if isinstance(other, Function):
            if self.subs is not None:
                for a, b in zip(self.subs, other.subs):
                    a.data[...] = b.data
            else:
                self.data[...] = other.data
        else:
            self.data[...] = _num(other)
        return self

    def interpolate(self, expr):
        n = self.mesh.n
        if self.space.kind == "VEC":
            assert isinstance(expr, VecExpr) and len(expr.items) == self.space.dim
            cols = [_nodal_values(e, n) for e in expr.items]      # all right-hand sides first: expr may read self
            for i, c in enumerate(cols):
                self.data[:, i] = c
        elif self.space.kind == "CG1":
            self.data[...] = _nodal_values(expr, n)
        else:
            raise NotImplementedError("stand-in: interpolate into a mixed function")
        return self

    # --- Cofunction side (a Cofunction of W.dual() is stored like a Function)
    def riesz_representation(self):
        """L2 Riesz map: the Function f with (f, v) = self(v) for all v, i.e. f = M^-1 self, component by component."""
        out = Function(self.space)
        for a, b in zip(out.subs, self.subs):
            a.data[...] = self.mesh.mass_solve(b.data)
        return out

    def __isub__(self, other):
        for a, b in zip(self.subs, other.subs):
            a.data -= b.data
        return self


Cofunction = Function


def _nodal_values(e, n):
    if isinstance(e, Nodal):
        return e.arr
    v = _num(e)
    if v is not None:
        return np.full(n, v)
    return Lin.of(e).evaluate(n)


def split(obj):
    if isinstance(obj, Function):
        assert obj.space.kind == "MIXED"
        return tuple(obj.subs)
    if isinstance(obj, (MixedTest, MixedTrial)):
        return obj.parts
    raise NotImplementedError(type(obj).__name__)


# ------------------------------------------------------------------------------------------- arguments and forms
class Test:
    def __init__(self, field, comp, grad=False):
        self.field, self.comp, self.grad = field, comp, grad


class _Indexable:
    def __init__(self, make, dim):
        self._make, self.dim = make, dim

    def __getitem__(self, i):
        return self._make(int(i))


class MixedTest:
    def __init__(self, space):
        self.space = space
        self.parts = tuple(_Indexable(lambda i, f=f: Test(f, i), space.subs[f].dim) for f in range(len(space.subs)))


class MixedTrial:
    def __init__(self, space):
        self.space = space
        self.parts = tuple(_Indexable(lambda i, f=f: Component(("trial", f, i)), space.subs[f].dim)
                           for f in range(len(space.subs)))


def TestFunction(space):
    if space.kind == "CG1":
        return Test("scalar", None)         # upstream's solve() builds (and never assembles) a few scalar check forms
    if space.kind != "MIXED":
        raise NotImplementedError("stand-in: test functions of the mixed or the scalar P1 space only")
    return MixedTest(space)


def TrialFunction(space):
    return MixedTrial(space)


def TrialFunctions(space):
    return MixedTrial(space).parts


def grad(e):
    if isinstance(e, Test):
        return Test(e.field, e.comp, True)
    l = Lin.of(e)
    if l.offset:
        raise NotImplementedError("stand-in: gradient of a constant offset")
    return Lin(l.terms, 0.0, True)


class Form:
    def __init__(self, terms=()):
        self.terms = list(terms)            # (coef, Lin, Test) | (coef, mixed Function, MixedTest)

    def __add__(self, o): return Form(self.terms + o.terms)
    def __sub__(self, o): return Form(self.terms + [(-c, l, t) for c, l, t in o.terms])
    def __iadd__(self, o):
        self.terms += o.terms
        return self
    def __isub__(self, o):
        self.terms += [(-c, l, t) for c, l, t in o.terms]
        return self
    def __neg__(self): return Form([(-c, l, t) for c, l, t in self.terms])
    def __rmul__(self, o):
        v = _num(o)
        return NotImplemented if v is None else Form([(v * c, l, t) for c, l, t in self.terms])
    __mul__ = __rmul__


class _Integrand:
    def __init__(self, pairs):
        self.pairs = pairs                  # [(Lin | mixed Function, Test | MixedTest)]

    def __mul__(self, measure):
        if measure is dx:
            return Form([(getattr(self, "scale", 1.0), l, t) for l, t in self.pairs])
        return self.__rmul__(measure)

    def __rmul__(self, o):                  # scalar * inner(...) [* dx]
        v = _num(o)
        if v is None:
            return NotImplemented
        out = _Integrand(self.pairs)
        out.scale = getattr(self, "scale", 1.0) * v
        return out


class _Measure:
    pass


dx = _Measure()


def inner(a, b):
    """(a, b) = integral of a conj(b); b is a (real) test function here, so the conjugate is void."""
    if isinstance(b, MixedTest):
        assert isinstance(a, Function) and a.space.kind == "MIXED"
        return _Integrand([(a, b)])
    if isinstance(b, _Indexable):           # whole vector-valued arguments: sum over the components
        return _Integrand([p for i in range(b.dim) for p in inner(a[i], b[i]).pairs])
    if not isinstance(b, Test):
        raise NotImplementedError("stand-in: inner(., test function) only")
    l = Lin.of(a)
    if l.grad != b.grad:
        raise NotImplementedError("stand-in: inner of a gradient with a value")
    return _Integrand([(l, b)])


def assemble(form, bcs=None):
    """Only what upstream's dead consistency check at :506-507 needs: the mass action on a mixed Function."""
    out = Function(form.terms[0][1].space)
    for c, f, t in form.terms:
        assert isinstance(t, MixedTest)
        for a, b in zip(out.subs, f.subs):
            a.data += c * (f.mesh.M @ b.data)
    return out


class DirichletBC:
    def __init__(self, subspace, value, where):
        assert where == "on_boundary"
        self.field = subspace.index if isinstance(subspace, SubSpace) else None


# --------------------------------------------------------------------------------------------------- assembly
def _layout(space):
    n = space.mesh.n
    dims = [s.dim for s in space.subs]
    offs = np.concatenate([[0], np.cumsum([n * d for d in dims])])
    return n, dims, offs


def assemble_affine(form, space, unknown, bcs):
    """Matrix A and vector b with  form(U, test) = A U - b  for every test function, U = `unknown` (a mixed Function,
    or the string 'trial').  Coefficient functions other than `unknown` contribute to b with their current data.
    Homogeneous Dirichlet rows / columns are replaced by the identity, b = 0 there."""
    mesh = space.mesh
    n, dims, offs = _layout(space)
    size = int(offs[-1])
    Mc, Kc = mesh.M.tocoo(), mesh.K.tocoo()
    rows, cols, vals = [], [], []
    b = np.zeros(size, dtype=complex)
    sub_of = {}
    if isinstance(unknown, Function):
        sub_of = {id(s): f for f, s in enumerate(unknown.subs)}
    for coef, lin, test in form.terms:
        S = Kc if test.grad else Mc
        r0 = int(offs[test.field])
        dr = dims[test.field]
        if lin.offset:
            raise NotImplementedError("stand-in: constant offset inside a form")
        for key, c in lin.terms.items():
            if key[0] == "trial" and unknown == "trial":
                f, comp = key[1], key[2]
            elif not isinstance(key[0], str) and id(key[0]) in sub_of:
                f, comp = sub_of[id(key[0])], key[1]
            elif isinstance(key[0], str):
                raise NotImplementedError("stand-in: trial function in a residual form")
            else:                                    # a known coefficient function: to the right-hand side
                vec = (mesh.K if test.grad else mesh.M) @ key[0].component(key[1])
                b[r0 + np.arange(n) * dr + test.comp] -= coef * c * vec
                continue
            rows.append(r0 + S.row * dr + test.comp)
            cols.append(int(offs[f]) + S.col * dims[f] + comp)
            vals.append(coef * c * S.data)
    if vals:
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(size, size),
                          dtype=complex).tocsr()
    else:
        A = sp.csr_matrix((size, size), dtype=complex)
    bc_idx = []
    for bc in bcs or []:
        f = bc.field
        for node in (0, n - 1):
            bc_idx.append(int(offs[f]) + node * dims[f] + np.arange(dims[f]))
    if bc_idx:
        bc_idx = np.concatenate(bc_idx)
        keep = np.ones(size)
        keep[bc_idx] = 0.0
        D = sp.diags(keep)
        A = D @ A @ D + sp.diags(1.0 - keep)
        b[bc_idx] = 0.0
    return A.tocsr(), b


class LinearVariationalProblem:
    def __init__(self, a, L, u, bcs=None):
        self.a, self.L, self.u, self.bcs = a, L, u, bcs or []


class LinearVariationalSolver:
    """a(u, v) = L(v): the matrix is assembled and factorised once, the right-hand side on every solve (its coefficient
    functions are read then)."""

    def __init__(self, problem, solver_parameters=None):
        self.problem = problem
        self._lu = None

    def solve(self):
        pr = self.problem
        space = pr.u.space
        if self._lu is None:
            A, _ = assemble_affine(pr.a, space, "trial", pr.bcs)
            self._lu = spla.splu(A.tocsc())
        _, minus_b = assemble_affine(pr.L, space, None, pr.bcs)      # L has coefficient functions only: "A U - b" = -L
        Dat._View(pr.u).setArray(self._lu.solve(-minus_b))


class NonlinearVariationalProblem:
    def __init__(self, F, u, bcs=None):
        self.F, self.u, self.bcs = F, u, bcs or []

    def affine_system(self):
        """The residual of this script is affine in U: F(U; v) = A U - b.  (``snes_type ksponly`` solves A U = b.)"""
        return assemble_affine(self.F, self.u.space, self.u, self.bcs)


class NonlinearVariationalSolver:
    """``snes_type ksponly`` from the zero initial guess: one linear solve of A U = b.  ``pc_type lu`` -> sparse LU
    (upstream: MUMPS).  ``pc_type python`` -> the class named by ``pc_python_type`` is instantiated, initialised and
    used as left preconditioner of the Krylov loop the driver script injected as ``ksp`` (PETSc's KSPGMRES is not
    runnable here; the driver passes its restatement together with PETSc's defaults for what the options leave open)."""
    python_pcs = {}        # class name -> class, registered by the driver script
    ksp = None             # callable(matvec, pc_apply, b, options) -> (x, iterations, history, reason)
    last = None            # (iterations, history, reason) of the latest Krylov solve

    def __init__(self, problem, solver_parameters=None):
        self.problem, self.params = problem, dict(solver_parameters or {})

    def solve(self):
        pr = self.problem
        A, b = pr.affine_system()
        if self.params.get("pc_type") == "python":
            cls = type(self).python_pcs[self.params["pc_python_type"].split(".")[-1]]
            pc = cls()
            pc.initialize(None)

            def pc_apply(x):
                xv, yv = HostVec(x), HostVec(np.zeros(b.size))
                pc.apply(None, xv, yv)
                return yv.array.copy()
            x, its, hist, reason = type(self).ksp(lambda z: A @ z, pc_apply, b, self.params)
            type(self).last = (its, hist, reason)
        else:
            x = spla.splu(A.tocsc()).solve(b)
        Dat._View(pr.u).setArray(x)


class PCBase:
    pass
