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
* UFL objects (``split``, ``TestFunctions``, ``dx``, ``ds``, ``grad``, ``pow``,
  ``sqrt``, ``derivative``): LIVE -- ``ufl_live.py`` builds a symbolic integrand
  from the reference's own form text (artery.py:191-230) and differentiates it
  (artery.py:238).
* ``NonlinearVariationalSolver.solve()``: a generic P1 / 3-point-Gauss assembler
  of whatever form it is given + DOLFIN's NewtonSolver loop (``ufl_live.py``).
  Generic FEniCS machinery is the only third-party behaviour restated; nothing
  in the golden generator imports the repo's oracle any more.
* ``XDMFFile`` / ``HDF5File``: capture what is written, in memory.

Not part of the product and not part of the test run on the GPU box; the
fixtures it produced are committed.
"""
import sys
import types

import numpy as np

import ufl_live

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


dx = ufl_live.Measure('dx')
ds = ufl_live.Measure('ds')


def sqrt(v):
    return ufl_live.usqrt(v) if isinstance(v, ufl_live.UExpr) else np.sqrt(v)


def pow(a, b):                                   # noqa: A001 (mirrors UFL's name)
    return a ** b if isinstance(a, ufl_live.UExpr) else np.power(a, b)


def grad(v):
    return ufl_live.Grad(v)


def derivative(F, U):
    return ufl_live.derivative(F, U)


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


class Expression(ufl_live.UExpr):
    """Pointwise-evaluated C++ string (the BC code calls ``a.f(x)``); inside a
    form it is a coefficient symbol that the assembler fills with the P2
    interpolant of these pointwise values."""

    def __init__(self, code, degree=None, **params):
        self._code = code
        self._names = list(params)
        self.degree = degree
        for k, v in params.items():
            setattr(self, k, v)

    @property
    def e(self):
        assert self.degree == 2, 'only degree=2 Expressions appear in the reference form'
        return ufl_live.expression_symbol(self)

    def _params(self):
        return {k: getattr(self, k) for k in self._names}

    def __call__(self, xval):
        if isinstance(self._code, (tuple, list)):
            return np.array([_evaluate(c, self._params(), xval) for c in self._code])
        return _evaluate(self._code, self._params(), xval)


class Function(object):
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

    def __getitem__(self, i):
        return ufl_live.coefficient_component(self, i)

    def split(self, deepcopy=False):
        out = []
        for c in range(self.V.dim):
            f = Function(FunctionSpace(self.V.mesh, FiniteElement('CG', 'interval', 1)))
            f.values[:, 0] = self.values[:, c]
            out.append(f)
        return tuple(out)


def split(u):
    return u[0], u[1]


def TestFunctions(V):
    return ufl_live.test_functions()


class DirichletBC(object):
    def __init__(self, V, value, where):
        self.V, self.value, self.where = V, value, where


class NonlinearVariationalProblem(object):
    def __init__(self, F, U, bcs, J=None):
        self.F, self.U, self.bcs = F, U, bcs
        self.J = J if J is not None else derivative(F, U)


# make_golden.py may install observer(problem, iterations, record) to log what a solve did;
# ``record_next`` = a dict makes the next solve store its residuals / Jacobians in it
solve_observer = None
record_next = None


class NonlinearVariationalSolver(object):
    """The reference's form and Jacobian form (artery.py:225-238), assembled
    and driven by ufl_live's generic assembler and NewtonSolver loop."""

    def __init__(self, problem):
        self.problem = problem
        self.parameters = _Params()

    def solve(self):
        global record_next
        p = self.problem
        max_it = self.parameters['newton_solver'].get('maximum_iterations', 50)
        record, record_next = record_next, None
        its = ufl_live.newton_solve(p.F, p.J, p.U, p.bcs, p.U.V.mesh, max_it=max_it, record=record)
        if solve_observer is not None:
            solve_observer(p, its)


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
