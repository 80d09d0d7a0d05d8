"""A small LIVE subset of UFL + DOLFIN assembly, used ONLY by make_golden.py.

Purpose: the reference builds its per-artery weak form as a UFL expression
(``/root/reference/arteryfe/artery.py:191-230``) and hands it, with
``derivative(F, U)`` (``artery.py:238``), to FEniCS.  FEniCS is absent here, so
this module makes those few UFL operations *live*: the unmodified reference
code constructs a symbolic integrand (sympy), ``derivative`` differentiates it
symbolically (sympy.diff, the Gateaux derivative UFL would form), and a generic
assembler integrates whatever came out on the interval mesh.  Nothing in here
knows the blood-flow equations: the terms, signs, DOLFIN_EPS placement, the
facet terms and the Jacobian all come from the reference's form text.

What is still a statement about third-party behaviour ([3P], FEniCS 2017.2.0,
not verifiable in this container) is confined to the GENERIC machinery and is
listed in DESIGN.md section 2:

* P1 Lagrange basis on ``IntervalMesh(Nx, 0, L)``, dofs ordered (A_j, q_j);
* ``Expression(..., degree=2)`` enters a form as its P2 Lagrange interpolant
  (vertex, vertex, midpoint values by exact pointwise evaluation);
* ``quadrature_degree = 4`` (``artery.py:9``) -> 3-point Gauss-Legendre on
  every cell; ``ds`` = evaluation at the two end vertices, weight 1, no normal
  unless the form writes one;
* ``DirichletBC``: residual row := U_dof - g, Jacobian row := identity row;
* ``NewtonSolver``: r = ||R||_2 over all rows, converged iff r/r0 < 1e-9 or
  r < 1e-10, tested before the first solve and after every update, count =
  linear solves, RuntimeError at ``maximum_iterations``;
* LU: dense LAPACK ``dgesv`` here (MUMPS there) -- round-off only.

Test infrastructure; never imported by the product.
"""
import numpy as np
import sympy as sp

# 3-point Gauss-Legendre on [0, 1] (FIAT Gauss-Jacobi(0,0), m = (4+2)//2 = 3)
GAUSS_XI = np.array([0.5 - 0.5 * np.sqrt(3.0 / 5.0), 0.5, 0.5 + 0.5 * np.sqrt(3.0 / 5.0)])
GAUSS_W = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])

_symbols = {}    # key -> sympy Symbol
_meaning = {}    # sympy Symbol -> key; key = (kind, object-or-number, component, derivative order)


def symbol_for(kind, obj, comp=0, deriv=0):
    key = (kind, id(obj) if kind in ('coef', 'expr') else obj, comp, deriv)
    if key not in _symbols:
        s = sp.Symbol('%s%d_%d_%d' % (kind, len(_symbols), comp, deriv), real=True)
        _symbols[key] = s
        _meaning[s] = (kind, obj, comp, deriv)
    return _symbols[key]


class UExpr(object):
    """A scalar UFL expression: wraps a sympy expression."""
    __array_ufunc__ = None          # numpy scalars defer to our reflected operators

    def __init__(self, e):
        self._e = e

    @property
    def e(self):
        return self._e

    @staticmethod
    def _lift(v):
        if isinstance(v, UExpr):
            return v.e
        if isinstance(v, (int, float, np.floating, np.integer)):
            # 17 significant digits: the generated numpy code then holds the exact double
            return sp.Float(float(v), 17) if not isinstance(v, (int, np.integer)) else sp.Integer(int(v))
        return None

    def _bin(self, other, fn):
        if isinstance(other, Measure):
            return NotImplemented
        o = self._lift(other)
        if o is None:
            return NotImplemented
        return UExpr(fn(self.e, o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return UExpr(-self.e)
    def __pos__(self): return self

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([(self.e, o.kind)])
        return self._bin(o, lambda a, b: a * b)


class Measure(object):
    __array_ufunc__ = None

    def __init__(self, kind):
        self.kind = kind

    def __rmul__(self, o):
        if isinstance(o, UExpr):
            return Form([(o.e, self.kind)])
        return Form([(UExpr._lift(o), self.kind)])


class Form(object):
    """Sum of integrals: [(sympy integrand, 'dx' | 'ds')]."""
    __array_ufunc__ = None

    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, o):
        return Form(self.integrals + o.integrals)

    def __sub__(self, o):
        return Form(self.integrals + [(-e, m) for e, m in o.integrals])

    def __neg__(self):
        return Form([(-e, m) for e, m in self.integrals])

    def __rmul__(self, c):
        c = UExpr._lift(c)
        return Form([(c * e, m) for e, m in self.integrals])

    __mul__ = __rmul__


class Grad(object):
    """grad(u): on an interval the only component is d/dx."""

    def __init__(self, u):
        self.u = u

    def __getitem__(self, i):
        assert i == 0
        e = self.u.e
        out = sp.Integer(0)
        for s in e.free_symbols:
            kind, obj, comp, deriv = _meaning[s]
            if kind == 'expr':
                raise NotImplementedError('grad of an Expression')
            if deriv == 0:          # P1: second derivatives vanish inside a cell
                out = out + sp.diff(e, s) * symbol_for(kind, obj, comp, 1)
        return UExpr(out)


def test_functions():
    return UExpr(symbol_for('test', 0, 0)), UExpr(symbol_for('test', 0, 1))


def coefficient_component(fn, comp):
    return UExpr(symbol_for('coef', fn, comp))


def expression_symbol(expr):
    return symbol_for('expr', expr)


def usqrt(v):
    return UExpr(sp.sqrt(v.e))


def upow(a, b):
    return a ** b


def derivative(F, U):
    """Gateaux derivative of the form F with respect to the Function U in the
    direction of a trial function of U's space (what ufl.derivative builds)."""
    cache = F.__dict__.setdefault('_derivatives', {})
    if id(U) in cache:                     # artery.py:238 re-derives every solve; the result is the same Form
        return cache[id(U)]
    out = []
    for e, m in F.integrals:
        de = sp.Integer(0)
        for s in e.free_symbols:
            kind, obj, comp, deriv = _meaning[s]
            if kind == 'coef' and obj is U:
                de = de + sp.diff(e, s) * symbol_for('trial', 0, comp, deriv)
        out.append((de, m))
    cache[id(U)] = Form(out)
    return cache[id(U)]


# ---------------------------------------------------------------------------
# generic assembly on IntervalMesh, P1 x P1, dof (node j, component c) -> 2j+c
# ---------------------------------------------------------------------------
class _Compiled(object):
    """One integral split by linearity into coefficient functions of the
    test (and trial) symbols, each lambdified over the remaining symbols."""

    def __init__(self, e, bilinear):
        tests = [symbol_for('test', 0, c, d) for c in (0, 1) for d in (0, 1)]
        trials = [symbol_for('trial', 0, c, d) for c in (0, 1) for d in (0, 1)]
        zero = {s: 0 for s in tests}
        assert sp.expand(e.subs(zero)) == 0, 'integrand is not linear in the test functions'
        self.terms = []      # (test comp, test deriv[, trial comp, trial deriv], callable, symbols)
        for c in (0, 1):
            for d in (0, 1):
                ce = sp.diff(e, symbol_for('test', 0, c, d))
                if ce == 0:
                    continue
                assert not (ce.free_symbols & set(tests))
                if not bilinear:
                    assert not (ce.free_symbols & set(trials))
                    self.terms.append(((c, d), self._fn(ce)))
                    continue
                assert sp.expand(ce.subs({s: 0 for s in trials})) == 0
                for c2 in (0, 1):
                    for d2 in (0, 1):
                        cc = sp.diff(ce, symbol_for('trial', 0, c2, d2))
                        if cc == 0:
                            continue
                        assert not (cc.free_symbols & set(trials))
                        self.terms.append(((c, d, c2, d2), self._fn(cc)))

    @staticmethod
    def _fn(e):
        syms = sorted(e.free_symbols, key=lambda s: s.name)
        return sp.lambdify(syms, e, modules='numpy', cse=True), syms


_compiled = {}
_p2_cache = {}


def _compile(form, bilinear):
    key = (id(form), bilinear)
    if key not in _compiled:
        _compiled[key] = (form, [(_Compiled(e, bilinear), m) for e, m in form.integrals])
    return _compiled[key][1]


def _cell_values(sym, mesh, cache=True):
    """Values of a non-argument symbol at the 3 Gauss points of every cell: (ncell, 3)."""
    kind, obj, comp, deriv = _meaning[sym]
    n, h = mesh.n, (mesh.b - mesh.a) / mesh.n
    xi = GAUSS_XI[None, :]
    if kind == 'coef':
        v = obj.values[:, comp]
        if deriv == 0:
            return v[:-1, None] * (1.0 - xi) + v[1:, None] * xi
        return np.repeat(((v[1:] - v[:-1]) / h)[:, None], 3, axis=1)
    if kind == 'expr':
        key = (id(obj), id(mesh))
        if key not in _p2_cache:
            x = mesh.coords
            vl = np.array([obj(t) for t in x[:-1]])
            vr = np.array([obj(t) for t in x[1:]])
            vm = np.array([obj(t) for t in 0.5 * (x[:-1] + x[1:])])
            n0 = (1 - xi) * (1 - 2 * xi)
            n1 = xi * (2 * xi - 1)
            nm = 4 * xi * (1 - xi)
            _p2_cache[key] = (obj, vl[:, None] * n0 + vr[:, None] * n1 + vm[:, None] * nm)
        return _p2_cache[key][1]
    raise ValueError(kind)


def _vertex_value(sym, mesh, node):
    kind, obj, comp, deriv = _meaning[sym]
    if deriv != 0:
        raise NotImplementedError('derivative in a facet integral')
    if kind == 'coef':
        return obj.values[node, comp]
    if kind == 'expr':
        return obj(mesh.coords[node])      # a P2 interpolant takes the exact value at a vertex
    raise ValueError(kind)


def assemble_vector(form, mesh):
    """Residual vector (Nx+1, 2) of a linear form."""
    n, h = mesh.n, (mesh.b - mesh.a) / mesh.n
    R = np.zeros((n + 1, 2))
    phi = (1.0 - GAUSS_XI, GAUSS_XI)
    dphi = (-1.0 / h, 1.0 / h)
    for comp_, m in _compile(form, False):
        for (c, d), (fn, syms) in comp_.terms:
            if m == 'dx':
                val = np.broadcast_to(fn(*[_cell_values(s, mesh) for s in syms]), (n, 3))
                for a in (0, 1):
                    basis = phi[a][None, :] if d == 0 else dphi[a]
                    R[a:n + a, c] += np.sum(val * basis * (GAUSS_W * h)[None, :], axis=1)
            else:
                assert d == 0, 'test-function derivative in a facet integral'
                for node in (0, n):
                    R[node, c] += fn(*[_vertex_value(s, mesh, node) for s in syms])
    return R


def assemble_matrix(form, mesh):
    """2x2-block tridiagonal matrix (Lo, Dg, Up), each (Nx+1, 2, 2):
    block[j][r, c] = d R[(j, r)] / d U[(j-1 | j | j+1, c)]."""
    n, h = mesh.n, (mesh.b - mesh.a) / mesh.n
    Lo = np.zeros((n + 1, 2, 2))
    Dg = np.zeros((n + 1, 2, 2))
    Up = np.zeros((n + 1, 2, 2))
    phi = (1.0 - GAUSS_XI, GAUSS_XI)
    dphi = (-1.0 / h, 1.0 / h)
    for comp_, m in _compile(form, True):
        for (c, d, c2, d2), (fn, syms) in comp_.terms:
            if m == 'dx':
                val = np.broadcast_to(fn(*[_cell_values(s, mesh) for s in syms]), (n, 3))
                for a in (0, 1):
                    ta = phi[a][None, :] if d == 0 else dphi[a]
                    for b in (0, 1):
                        tb = phi[b][None, :] if d2 == 0 else dphi[b]
                        blk = np.sum(val * ta * tb * (GAUSS_W * h)[None, :], axis=1)
                        if a == b:
                            Dg[a:n + a, c, c2] += blk
                        elif a == 0:
                            Up[0:n, c, c2] += blk
                        else:
                            Lo[1:n + 1, c, c2] += blk
            else:
                assert d == 0 and d2 == 0
                for node in (0, n):
                    Dg[node, c, c2] += fn(*[_vertex_value(s, mesh, node) for s in syms])
    return Lo, Dg, Up


def dirichlet_rows(bcs, mesh):
    """[(node, component, value)] of a list of stand-in DirichletBC objects."""
    rows = []
    for bc in bcs:
        for node in (0, mesh.n):
            if not bc.where([mesh.coords[node]], True):
                continue
            g = bc.value(mesh.coords[node])
            comp = getattr(bc.V, 'component', None)
            if comp is None:
                rows += [(node, 0, float(g[0])), (node, 1, float(g[1]))]
            else:
                rows.append((node, comp, float(g)))
    return rows


def apply_bc_vector(R, U, rows):
    R = R.copy()
    for node, comp, g in rows:
        R[node, comp] = U[node, comp] - g
    return R


def apply_bc_matrix(blocks, rows):
    Lo, Dg, Up = (b.copy() for b in blocks)
    for node, comp, g in rows:
        Lo[node, comp, :] = 0
        Up[node, comp, :] = 0
        Dg[node, comp, :] = 0
        Dg[node, comp, comp] = 1
    return Lo, Dg, Up


def dense(blocks):
    Lo, Dg, Up = blocks
    n1 = Dg.shape[0]
    M = np.zeros((2 * n1, 2 * n1))
    for j in range(n1):
        M[2 * j:2 * j + 2, 2 * j:2 * j + 2] = Dg[j]
        if j > 0:
            M[2 * j:2 * j + 2, 2 * j - 2:2 * j] = Lo[j]
        if j < n1 - 1:
            M[2 * j:2 * j + 2, 2 * j + 2:2 * j + 4] = Up[j]
    return M


def newton_solve(F, J, U, bcs, mesh, rtol=1e-9, atol=1e-10, max_it=50, record=None):
    """dolfin::NewtonSolver::solve with the residual criterion ([3P], see the
    module docstring).  ``record`` (a dict) receives the residual / Jacobian of
    the first two iterates for the fem_* fixtures."""
    rows = dirichlet_rows(bcs, mesh)
    Rraw = assemble_vector(F, mesh)
    R = apply_bc_vector(Rraw, U.values, rows)
    r0 = r = np.sqrt(np.sum(R * R))
    it = 0
    if record is not None:
        record.update(bc_rows=np.array(rows), U0=U.values.copy(), R0_raw=Rraw, R0=R)
    while not (r / r0 < rtol or r < atol):
        if it >= max_it or not np.isfinite(r):
            raise RuntimeError('Newton solver did not converge (it=%d, r=%g)' % (it, r))
        Jraw = assemble_matrix(J, mesh)
        Jbc = apply_bc_matrix(Jraw, rows)
        dU = np.linalg.solve(dense(Jbc), R.reshape(-1)).reshape(-1, 2)
        if record is not None and it < 2:
            record.update({'J%d_raw' % it: np.array(Jraw), 'J%d' % it: np.array(Jbc), 'dU%d' % it: dU})
        U.values -= dU
        it += 1
        Rraw = assemble_vector(F, mesh)
        R = apply_bc_vector(Rraw, U.values, rows)
        r = np.sqrt(np.sum(R * R))
        if record is not None and it == 1:
            record.update(U1=U.values.copy(), R1_raw=Rraw, R1=R)
    if record is not None:
        record.update(its=it, U_final=U.values.copy(), r_final=r)
    return it
