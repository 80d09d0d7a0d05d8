"""Minimal stand-in for the ``dolfin`` module, used ONLY by make_golden.py.

It lets the UNMODIFIED reference modules (``/root/reference/arteryfe/*.py``)
be imported and driven in a container that has no FEniCS, so that the
reference's own Python code for the hot path -- the time loop, set_bcs,
the 18-unknown junction Newton, the Windkessel Picard, compute_U_half, flux,
source, CFL_term, adjust_dex, initial conditions, parameter handling -- produces
the golden vectors under ``tests/golden/``.

What the stand-in provides:

* ``Expression``: the C++ strings of artery.py are evaluated pointwise with
  numpy (``pow``/``exp``/``sqrt``/``pi``/``x[0]``), nested Expressions and
  Functions included; keyword parameters stay settable attributes.
* ``Function``: nodal values on ``IntervalMesh(Nx, 0, L)``; ``assign`` evaluates
  an Expression at the vertices; calling it is P1 interpolation.
* UFL objects (``split``, ``TestFunctions``, ``dx``, ``ds``, ``grad`` ...): inert
  symbols -- the variational form cannot be compiled here.
* ``NonlinearVariationalSolver.solve()``: THE ONE SUBSTITUTION.  FEniCS' assembly
  + Newton + LU is replaced by a callback (make_golden.py installs the numpy
  oracle's per-artery solve).  Everything else that runs is reference code.
* ``XDMFFile`` / ``HDF5File``: capture what is written, in memory.

Not part of the product and not part of the test run on the GPU box; the
fixtures it produced are committed.
"""
import sys
import types

import numpy as np

DOLFIN_EPS = 3.0e-16
pi = np.pi


class _Params(dict):
    def __getitem__(self, k):
        if k not in self:
            self[k] = _Params()
        return dict.__getitem__(self, k)


parameters = _Params()


class _Comm(object):
    def tompi4py(self):
        return self


def mpi_comm_world():
    return _Comm()


def set_log_level(level):
    pass


def near(a, b, tol=3.0e-16):
    return abs(a - b) < tol


class Sym(object):
    """Inert UFL symbol: absorbs arithmetic."""

    def _op(self, *a, **k):
        return Sym()

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = _op
    __truediv__ = __rtruediv__ = __pow__ = __rpow__ = __neg__ = __pos__ = _op
    __getitem__ = __call__ = _op


dx = Sym()
ds = Sym()


def sqrt(v):
    return Sym() if isinstance(v, (Sym, Function, Expression)) else np.sqrt(v)


def pow(a, b):                                   # noqa: A001 (mirrors UFL's name)
    return Sym() if isinstance(a, (Sym, Function, Expression)) else np.power(a, b)


def grad(v):
    return Sym()


def derivative(F, U):
    return Sym()


class IntervalMesh(object):
    def __init__(self, n, a, b):
        self.n, self.a, self.b = n, a, b
        self.coords = a + (b - a) / n * np.arange(n + 1)

    def ufl_cell(self):
        return 'interval'


class FiniteElement(object):
    def __init__(self, family, cell, degree):
        self.dim = 1

    def __mul__(self, other):
        e = FiniteElement('CG', 'interval', 1)
        e.dim = self.dim + other.dim
        return e


class FunctionSpace(object):
    def __init__(self, mesh, element, degree=None):
        self.mesh = mesh
        self.dim = element.dim if isinstance(element, FiniteElement) else 1

    def sub(self, i):
        s = FunctionSpace(self.mesh, FiniteElement('CG', 'interval', 1))
        s.component = i
        return s


def _evaluate(code, params, xval):
    ns = {'pow': np.power, 'exp': np.exp, 'sqrt': np.sqrt, 'log': np.log,
          'pi': np.pi, 'x': [xval]}
    for k, v in params.items():
        ns[k] = v(xval) if callable(v) else v
    return eval(code, {'__builtins__': {}}, ns)


class Expression(Sym):
    def __init__(self, code, degree=None, **params):
        object.__setattr__(self, '_code', code)
        object.__setattr__(self, '_names', list(params))
        for k, v in params.items():
            object.__setattr__(self, k, v)

    def _params(self):
        return {k: getattr(self, k) for k in self._names}

    def __call__(self, xval):
        if isinstance(self._code, (tuple, list)):
            return np.array([_evaluate(c, self._params(), xval) for c in self._code])
        return _evaluate(self._code, self._params(), xval)


class Function(Sym):
    def __init__(self, V):
        self.V = V
        self.values = np.zeros((V.mesh.n + 1, V.dim))

    def set_allow_extrapolation(self, flag):
        pass

    def assign(self, other):
        if isinstance(other, Function):
            self.values[:] = other.values
        else:
            for j, xj in enumerate(self.V.mesh.coords):
                self.values[j] = other(xj)

    def __call__(self, xval):
        m = self.V.mesh
        h = (m.b - m.a) / m.n
        j = min(max(int(np.floor((xval - m.a) / h)), 0), m.n - 1)
        xi = (xval - (m.a + j * h)) / h
        v = self.values[j] * (1.0 - xi) + self.values[j + 1] * xi
        return v if self.V.dim > 1 else float(v[0])

    def split(self, deepcopy=False):
        out = []
        for c in range(self.V.dim):
            f = Function(FunctionSpace(self.V.mesh, FiniteElement('CG', 'interval', 1)))
            f.values[:, 0] = self.values[:, c]
            out.append(f)
        return tuple(out)


def split(u):
    return Sym(), Sym()


def TestFunctions(V):
    return Sym(), Sym()


class DirichletBC(object):
    def __init__(self, V, value, where):
        self.V, self.value, self.where = V, value, where


class NonlinearVariationalProblem(object):
    def __init__(self, F, U, bcs, J=None):
        self.U, self.bcs = U, bcs


# make_golden.py installs: callback(problem) -> None, updating problem.U.values
artery_solve_callback = None


class NonlinearVariationalSolver(object):
    def __init__(self, problem):
        self.problem = problem
        self.parameters = _Params()

    def solve(self):
        artery_solve_callback(self.problem)


class HDF5File(object):
    def __init__(self, comm, name, mode):
        pass

    def write(self, obj, name):
        pass

    def close(self):
        pass


class XDMFFile(object):
    written = {}

    def __init__(self, *args):
        self.name = args[-1]
        XDMFFile.written.setdefault(self.name, [])

    def write_checkpoint(self, u, label, t):
        XDMFFile.written[self.name].append((t, u.values[:, 0].copy()))


def install():
    """Register this module as ``dolfin`` and stub the other absent imports of
    arteryfe/utils.py (matplotlib, SafeConfigParser)."""
    import configparser
    me = sys.modules[__name__]
    sys.modules['dolfin'] = me
    if not hasattr(configparser, 'SafeConfigParser'):
        configparser.SafeConfigParser = configparser.ConfigParser
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


__all__ = [n for n in dir() if not n.startswith('_') and n not in ('sys', 'types', 'np')]
