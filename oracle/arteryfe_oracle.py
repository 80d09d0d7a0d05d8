"""numpy restatement of artery.fe's per-timestep network solve (the ORACLE).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Never imported by the
product package ``bloodflow_b200``.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Two kinds of semantics are restated:

* [REF]  the reference's own numpy code for the boundary conditions
  (``arteryfe/artery_network.py``) -- pinned against the reference itself by
  ``tests/golden/make_golden.py`` / ``tests/test_oracle_golden.py``;
* [3P]   FEniCS 2017.2.0 semantics (P1 point evaluation, degree-2 coefficient
  interpolation, degree-4 Gauss quadrature, DirichletBC rows, NewtonSolver
  control) which the reference calls but does not contain -- parity unpinned.

The code is deliberately plain: formulas are written the way the reference
writes them (``sqrt(A0*A)``, ``x**(3/2)``, ``dt/R1/R2/CT`` ...) so that the
oracle differs from the reference arithmetic only where FEniCS' generated code
(``-ffast-math``) makes bitwise agreement impossible anyway.
"""
import configparser
import copy

import numpy as np
from scipy.interpolate import interp1d
from scipy.linalg import solve_banded

# [3P] dolfin/common/constants.h: DOLFIN_EPS
DOLFIN_EPS = 3.0e-16

# [3P] FIAT Gauss-Jacobi(0,0), 3 points, on the reference interval [0, 1]:
# parameters["form_compiler"]["quadrature_degree"] = 4   (artery.py:9)
_S35 = np.sqrt(3.0 / 5.0)
GAUSS_XI = np.array([0.5 - 0.5 * _S35, 0.5, 0.5 + 0.5 * _S35])
GAUSS_W = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])


# ---------------------------------------------------------------------------
# host set-up  (param_parser.py, utils.py)
# ---------------------------------------------------------------------------
def load_cfg(path):
    """param_parser.py:26-81 -- three sections, values eval'd, comma => array.

    ``[Parameters]``: a value containing a comma becomes a float array,
    anything else is ``eval``'d (param_parser.py:59-63).  The other two
    sections are ``eval``'d with a string fallback (param_parser.py:76-80).
    """
    cp = configparser.ConfigParser()
    cp.optionxform = str
    if not cp.read(path):
        raise OSError("cannot read %s" % path)
    param = {}
    for key, value in cp.items('Parameters'):
        if ',' in value:
            param[key] = np.array([float(v) for v in value.split(',')])
        else:
            param[key] = eval(value)
    out = []
    for section in ('Geometry', 'Solution'):
        d = {}
        for key, value in cp.items(section):
            try:
                d[key] = eval(value)
            except Exception:
                d[key] = value
        out.append(d)
    return param, out[0], out[1]


def build_geometry(order, Ru, Rd, alpha, L, R_term):
    """artery_network.py:195-217 -- geometric-series tree.

    Daughter radii are ``alpha * parent * multiplier``; daughters on the last
    level get ``Rd = R_term``; lengths are ``L * Ru`` and the second daughter is
    shortened by 0.01.  Operates on (and returns) copies.
    """
    N = 2 ** order - 1
    Ru = np.array(Ru, dtype=float)
    Rd = np.array(Rd, dtype=float)
    Ll = np.zeros(N)
    Ll[0] = L * Ru[0]
    first, stop = 0, 0
    for level in range(order):
        for p in range(first, stop):
            d1, d2 = 2 * p + 1, 2 * p + 2
            Ru[d1] = alpha * Ru[p] * Ru[d1]
            Ru[d2] = max(0, alpha * Ru[p]) * Ru[d2]
            if 2 ** (level + 1) - 1 < len(Rd):
                Rd[d1] = alpha * Rd[p] * Rd[d1]
                Rd[d2] = max(0, alpha * Rd[p]) * Rd[d2]
            else:
                Rd[d1] = Rd[d2] = R_term
            Ll[d1] = L * Ru[d1]
            Ll[d2] = L * Ru[d2] - 0.01
        first += int(2 ** (level - 1))
        stop += int(2 ** level)
    return Ru, Rd, Ll


def nondimensionalise_parameters(param):
    """utils.py:47-104.  ``R_term`` / ``p_term`` are optional here."""
    nd = copy.deepcopy(param)
    rc, rho, qc = param['rc'], param['rho'], param['qc']
    nd['Ru'] = param['Ru'] / rc
    nd['Rd'] = param['Rd'] / rc
    if 'R_term' in param:
        nd['R_term'] = param['R_term'] / rc
    nd['L'] = param['L'] / rc
    nd['k1'] = param['k1'] * rc ** 4 / rho / qc ** 2
    nd['k2'] = param['k2'] * rc
    nd['k3'] = param['k3'] * rc ** 4 / rho / qc ** 2
    nd['Re'] = param['qc'] / param['nu'] / rc
    nd['nu'] = param['nu'] * rc / qc
    nd['p0'] = param['p0'] * rc ** 4 / rho / qc ** 2
    if 'p_term' in param:
        nd['p_term'] = param['p_term'] * rc ** 4 / rho / qc ** 2
    nd['R1'] = param['R1'] * rc ** 4 / rho / qc
    nd['R2'] = param['R2'] * rc ** 4 / rho / qc
    nd['CT'] = param['CT'] * rho * qc ** 2 / rc ** 7
    return nd


def read_inlet(data_location, Nt):
    """utils.py:173-200 -- linear interpolation of the CSV at linspace(0,T,Nt)."""
    data_q = np.genfromtxt(data_location, delimiter=',')
    tt, qq = data_q[:, 0], data_q[:, 1]
    T = data_q[-1, 0]
    q_ins = interp1d(tt, qq)(np.linspace(0, T, Nt))
    return T, q_ins


# ---------------------------------------------------------------------------
# one vessel  (artery.py)
# ---------------------------------------------------------------------------
class P1Function(object):
    """[3P] dolfin.Function on a P1 (vector) space over IntervalMesh(Nx, 0, L).

    ``values`` has one row per vertex, vertices ``x_j = j*dx``.  Calling it at a
    point is P1 interpolation in the containing cell (artery.py:155 allows
    extrapolation; the last/first cell is extended).
    """

    def __init__(self, dx, values):
        self.dx = dx
        self.values = values

    def __call__(self, x):
        n_cells = self.values.shape[0] - 1
        j = int(np.floor(x / self.dx))
        j = min(max(j, 0), n_cells - 1)
        xi = (x - j * self.dx) / self.dx
        return self.values[j] * (1.0 - xi) + self.values[j + 1] * xi


class OracleArtery(object):
    """Restates arteryfe.artery.Artery (artery.py:15-415)."""

    def __init__(self, i, T, param, root=False, leaf=False):
        # artery.py:56-64
        self.i, self.root, self.leaf, self.T = i, root, leaf, T
        self.param = copy.deepcopy(param)
        self.param['Ru'] = param['Ru'][i]
        self.param['Rd'] = param['Rd'][i]
        self.param['L'] = param['L'][i]
        self.newton_its = 0

    # -- Expressions, artery.py:110-118 (exact, pointwise) -------------------
    def r0(self, x):
        return self.Ru * np.power(self.Rd / self.Ru, x / self.L)

    def A0(self, x):
        return np.pi * np.power(self.r0(x), 2)

    def f(self, x):
        return 4.0 / 3.0 * (self.k1 * np.exp(self.k2 * self.r0(x)) + self.k3)

    def dfdr(self, x):
        return 4.0 / 3.0 * self.k1 * self.k2 * np.exp(self.k2 * self.r0(x))

    def drdx(self, x):
        return self.logRdRu / self.L * self.r0(x)

    def define_geometry(self, geo):
        # artery.py:67-118
        p = self.param
        self.Nx = geo['Nx']
        self.L, self.Ru, self.Rd = p['L'], p['Ru'], p['Rd']
        self.k1, self.k2, self.k3 = p['k1'], p['k2'], p['k3']
        self.dx = self.L / self.Nx
        self.dt = self.T / geo['Nt']
        self.dex = self.dx
        self.db = np.sqrt(p['nu'] * self.T / 2 / np.pi)
        if self.leaf and self.Rd == 1:
            self.Rd = p['R_term']
        self.logRdRu = np.log(self.Rd / self.Ru)
        # [3P] IntervalMesh vertex coordinates and P2 dof points
        self.xn = self.dx * np.arange(self.Nx + 1)
        xm = 0.5 * (self.xn[:-1] + self.xn[1:])
        # [3P] degree=2 Expressions enter the form as their P2 interpolant:
        # values at (left vertex, right vertex, midpoint) of every cell, then
        # evaluated at the Gauss points with the P2 Lagrange basis.
        xi = GAUSS_XI
        n0 = (1 - xi) * (1 - 2 * xi)
        n1 = xi * (2 * xi - 1)
        nm = 4 * xi * (1 - xi)
        self.qp = {}
        for name in ('f', 'A0', 'dfdr', 'drdx'):
            fn = getattr(self, name)
            vl, vr, vm = fn(self.xn[:-1]), fn(self.xn[1:]), fn(xm)
            self.qp[name] = (vl[:, None] * n0[None, :] + vr[:, None] * n1[None, :]
                             + vm[:, None] * nm[None, :])

    def define_solution(self, q0, theta=0.5, bc_tol=1.e-14):
        # artery.py:121-189
        self.q0, self.theta = q0, theta
        Un = np.empty((self.Nx + 1, 2))
        Un[:, 0] = self.A0(self.xn)
        Un[:, 1] = q0
        self._Un = Un
        self._U = Un.copy()
        self.pn_values = np.full(self.Nx + 1, self.param['p0'])
        if self.root:
            self.q_in = q0
        else:
            self.U_in = np.array([self.A0(0), q0])
        if self.leaf:
            self.A_out = self.A0(self.L)
        else:
            self.U_out = np.array([self.A0(self.L), q0])

    @property
    def Un(self):
        return P1Function(self.dx, self._Un)

    @property
    def U(self):
        return P1Function(self.dx, self._U)

    # -- weak form, artery.py:191-230 + [3P] assembly ------------------------
    def _F2(self, A, q, f, A0):
        return np.power(q, 2) / A + f * np.sqrt(A0 * A)

    def _S2(self, A, q, f, A0, dfdr, drdx):
        Re = self.param['Re']
        return (-2 * np.sqrt(np.pi) / self.db / Re * q / np.sqrt(A)
                + (2 * np.sqrt(A) * (np.sqrt(np.pi) * f + np.sqrt(A0) * dfdr)
                   - A * dfdr) * drdx)

    def assemble(self, U, jacobian=False, apply_bc=True):
        """Residual (Nx+1, 2) and, optionally, the 2x2-block tridiagonal
        Jacobian (lower, diag, upper: each (Nx+1, 2, 2); block[r, c] is
        d R[(node, r)] / d U[(neighbour, c)]) of the theta-scheme weak form,
        Dirichlet rows applied the DOLFIN way."""
        n, h, dt, th = self.Nx, self.dx, self.dt, self.theta
        Re = self.param['Re']
        rpi = np.sqrt(np.pi)
        Kr = 2 * rpi / self.db / Re
        Al, Ar, ql, qr = U[:-1, 0], U[1:, 0], U[:-1, 1], U[1:, 1]
        Anl, Anr = self._Un[:-1, 0], self._Un[1:, 0]
        qnl, qnr = self._Un[:-1, 1], self._Un[1:, 1]
        R = np.zeros((n + 1, 2))
        if jacobian:
            Lo = np.zeros((n + 1, 2, 2))
            Dg = np.zeros((n + 1, 2, 2))
            Up = np.zeros((n + 1, 2, 2))
        qx = (qr - ql) / h
        qnx = (qnr - qnl) / h
        for g in range(3):
            xi, w = GAUSS_XI[g], GAUSS_W[g] * h
            phi = (1.0 - xi, xi)
            dphi = (-1.0 / h, 1.0 / h)
            fq, A0q = self.qp['f'][:, g], self.qp['A0'][:, g]
            dfq, drq = self.qp['dfdr'][:, g], self.qp['drdx'][:, g]
            A = Al * phi[0] + Ar * phi[1]
            Ae = A + DOLFIN_EPS
            q = ql * phi[0] + qr * phi[1]
            An = Anl * phi[0] + Anr * phi[1]
            qn = qnl * phi[0] + qnr * phi[1]
            F2 = th * self._F2(Ae, q, fq, A0q) + (1 - th) * self._F2(An, qn, fq, A0q)
            S2 = (th * self._S2(Ae, q, fq, A0q, dfq, drq)
                  + (1 - th) * self._S2(An, qn, fq, A0q, dfq, drq))
            if jacobian:
                dF2dA = -np.power(q, 2) / np.power(Ae, 2) + fq * A0q / (2 * np.sqrt(A0q * Ae))
                dF2dq = 2 * q / Ae
                dS2dA = (Kr / 2 * q / (Ae * np.sqrt(Ae))
                         + ((rpi * fq + np.sqrt(A0q) * dfq) / np.sqrt(Ae) - dfq) * drq)
                dS2dq = -Kr / np.sqrt(Ae)
            for i in (0, 1):
                R[i:n + i, 0] += w * ((A - An) * phi[i] + dt * (th * qx + (1 - th) * qnx) * phi[i])
                R[i:n + i, 1] += w * ((q - qn) * phi[i] - dt * F2 * dphi[i] - dt * S2 * phi[i])
                if not jacobian:
                    continue
                for j in (0, 1):
                    blk = np.zeros((n, 2, 2))
                    blk[:, 0, 0] = w * phi[i] * phi[j]
                    blk[:, 0, 1] = w * dt * th * dphi[j] * phi[i]
                    blk[:, 1, 0] = -w * dt * th * (dF2dA * phi[j] * dphi[i] + dS2dA * phi[j] * phi[i])
                    blk[:, 1, 1] = w * (phi[i] * phi[j]
                                        - dt * th * (dF2dq * phi[j] * dphi[i] + dS2dq * phi[j] * phi[i]))
                    if i == j:
                        Dg[i:n + i] += blk
                    elif i == 0:
                        Up[0:n] += blk
                    else:
                        Lo[1:n + 1] += blk
        # exterior-facet terms: "+ F2*v2*ds" at BOTH ends, no normal
        # (artery.py:197-198, 206-207).  Vertex values of a P2 interpolant are
        # the exact Expression values.
        for node, x in ((0, 0.0), (n, self.xn[n])):
            fv, A0v = self.f(x), self.A0(x)
            Ae = U[node, 0] + DOLFIN_EPS
            q = U[node, 1]
            R[node, 1] += dt * (th * self._F2(Ae, q, fv, A0v)
                                + (1 - th) * self._F2(self._Un[node, 0], self._Un[node, 1], fv, A0v))
            if jacobian:
                Dg[node, 1, 0] += dt * th * (-np.power(q, 2) / np.power(Ae, 2)
                                             + fv * A0v / (2 * np.sqrt(A0v * Ae)))
                Dg[node, 1, 1] += dt * th * 2 * q / Ae
        # [3P] DirichletBC: residual row := U_dof - g ; Jacobian row := identity
        for node, comp, g in (self._dirichlet() if apply_bc else ()):
            R[node, comp] = U[node, comp] - g
            if jacobian:
                Lo[node, comp, :] = 0
                Up[node, comp, :] = 0
                Dg[node, comp, :] = 0
                Dg[node, comp, comp] = 1
        if jacobian:
            return R, (Lo, Dg, Up)
        return R

    def _dirichlet(self):
        # artery.py:171-189
        n = self.Nx
        if self.root:
            rows = [(0, 1, self.q_in)]
        else:
            rows = [(0, 0, self.U_in[0]), (0, 1, self.U_in[1])]
        if self.leaf:
            rows.append((n, 0, self.A_out))
        else:
            rows += [(n, 0, self.U_out[0]), (n, 1, self.U_out[1])]
        return rows

    @staticmethod
    def _block_tridiag_solve(blocks, R):
        """LU with partial pivoting on the banded matrix (LAPACK dgbsv), the
        nearest stand-in for the reference's MUMPS LU (artery.py:241)."""
        Lo, Dg, Up = blocks
        m = 2 * Dg.shape[0]
        ab = np.zeros((7, m))
        for r in (0, 1):
            for c in (0, 1):
                rows = 2 * np.arange(Dg.shape[0]) + r
                for blk, off in ((Lo, -2), (Dg, 0), (Up, 2)):
                    cols = rows - r + off + c
                    ok = (cols >= 0) & (cols < m)
                    ab[3 + rows[ok] - cols[ok], cols[ok]] = blk[ok, r, c]
        return solve_banded((3, 3), ab, R.reshape(-1)).reshape(-1, 2)

    def solve(self, rtol=1e-9, atol=1e-10, max_it=1000):
        """artery.py:233-244 + [3P] dolfin::NewtonSolver::solve.

        Initial guess is the persisting ``U``; the test is applied before the
        first solve and after every update; the count is the number of linear
        solves.  Non-convergence raises RuntimeError like DOLFIN does."""
        U = self._U
        R = self.assemble(U)
        r0 = r = np.sqrt(np.sum(R * R))
        it = 0
        while not (r / r0 < rtol or r < atol):
            if it >= max_it or not np.isfinite(r):
                self.newton_its = it
                raise RuntimeError("artery %d: Newton solver did not converge "
                                   "(it=%d, r=%g)" % (self.i, it, r))
            R, blocks = self.assemble(U, jacobian=True)
            U -= self._block_tridiag_solve(blocks, R)
            it += 1
            R = self.assemble(U)
            r = np.sqrt(np.sum(R * R))
        self.newton_its = it
        self.last_residual = r

    def update_solution(self):
        # artery.py:247-251
        self._Un[:] = self._U

    def update_pressure(self):
        # artery.py:254-261: nodal interpolation of p0 + f*(1-sqrt(A0/A))
        x = self.xn
        self.pn_values = self.param['p0'] + self.f(x) * (1 - np.sqrt(self.A0(x) / self._Un[:, 0]))

    def compute_outlet_pressure(self, A):
        # artery.py:285-301
        return self.param['p0'] + self.f(self.L) * (1 - np.sqrt(self.A0(self.L) / A))

    def CFL_term(self, x, A, q):
        # artery.py:304-325 (dimensional rho on purpose)
        rho = self.param['rho']
        c = np.sqrt(self.f(x) / 2 / rho * np.sqrt(self.A0(x) / A))
        return 1 / max(abs(q / A + c), abs(q / A - c))

    def check_CFL(self, x, A, q):
        # artery.py:328-346
        return self.dt / self.dex < self.CFL_term(x, A, q)

    def adjust_dex(self, x, A, q, margin=0.05):
        # artery.py:349-365
        self.dex = (1 + margin) * self.dt / self.CFL_term(x, A, q)


# ---------------------------------------------------------------------------
# the network  (artery_network.py)
# ---------------------------------------------------------------------------
class OracleNetwork(object):
    """Restates arteryfe.artery_network.ArteryNetwork (artery_network.py:14-946).

    ``wk_scale`` are the two hard-coded Windkessel factors of HEAD
    (artery_network.py:408-409); the default reproduces HEAD.
    """

    def __init__(self, param, geo, sol, wk_scale=(0.3, 0.01)):
        # artery_network.py:56-130
        param = copy.deepcopy(param)
        order = param['order']
        self.N = 2 ** order - 1
        self.wk_scale = wk_scale
        if 'alpha' in param:
            if 'R_term' not in param:
                raise KeyError('R_term')
            param['Ru'], param['Rd'], param['L'] = build_geometry(
                order, param['Ru'], param['Rd'], param['alpha'], param['L'], param['R_term'])
        n_leaf_max = 2 ** (order - 1)
        for key in ('R1', 'R2', 'CT'):           # scalar => every leaf
            if np.ndim(param[key]) == 0:
                param[key] = np.full(n_leaf_max, float(param[key]))
        self.nondim = nondimensionalise_parameters(param)
        self.geo, self.sol = dict(geo), dict(sol)
        rc, qc = self.nondim['rc'], self.nondim['qc']
        Nt = self.geo['Nt']
        self.T, self.q_ins = read_inlet(self.sol['inlet_flow_location'], Nt)
        self.T = self.T * qc / rc ** 3
        self.q_ins = self.q_ins / qc
        self.dt = self.T / Nt
        for key in ('Ru', 'Rd', 'L'):
            assert len(self.nondim[key]) == self.N

        self.arteries = [None] * self.N
        for i in range(self.N):
            if self.nondim['Ru'][i] != 0:
                self.arteries[i] = OracleArtery(i, self.T, self.nondim)
        self.arteries[0].root = True
        self.range_parent_arteries = []
        self.range_daughter_arteries = []
        self.range_leaf_arteries = []
        for i in range(self.N):
            if i == 0:
                self.range_parent_arteries.append(i)
            elif self.arteries[i] is not None:
                self.range_daughter_arteries.append(i)
                d1, d2 = self.daughter_arteries(i)
                if d1 is None and d2 is None:
                    self.arteries[i].leaf = True
                    self.range_leaf_arteries.append(i)
                else:
                    self.range_parent_arteries.append(i)
        for j, i in enumerate(self.range_leaf_arteries):
            for key in ('R1', 'R2', 'CT'):
                self.arteries[i].param[key] = self.nondim[key][j]
        self.define_geometry()
        self.define_solution()
        self.junction_solves = np.zeros(len(self.range_parent_arteries), dtype=int)
        self.picard_its = np.zeros(len(self.range_leaf_arteries), dtype=int)

    # -- topology, artery_network.py:133-192 ---------------------------------
    def daughter_arteries(self, i):
        d = []
        for k in (2 * i + 1, 2 * i + 2):
            d.append(None if k > self.N - 1 or self.arteries[k] is None else k)
        return d[0], d[1]

    def parent_artery(self, i):
        return (i - 1) // 2

    def sister_artery(self, i):
        return i - 1 if i % 2 == 0 else i + 1

    def define_geometry(self):
        # artery_network.py:244-255
        for a in self.arteries:
            if a is not None:
                a.define_geometry(self.geo)

    def define_solution(self):
        # artery_network.py:258-279
        theta = self.sol['theta']
        self.arteries[0].define_solution(self.q_ins[0], theta)
        for i in self.range_daughter_arteries:
            a, s = self.arteries[i], self.arteries[self.sister_artery(i)]
            p = self.arteries[self.parent_artery(i)]
            q0 = a.A0(0) / (a.A0(0) + s.A0(0)) * p.q0
            a.define_solution(q0, theta)
        self.x = np.ones([len(self.range_parent_arteries), 18])

    # -- boundary physics ------------------------------------------------------
    def flux(self, a, U, x):
        # artery_network.py:282-300 -- note: q**2, not q**2/A
        return np.array([U[1], U[1] ** 2 + a.f(x) * np.sqrt(a.A0(x) * U[0])])

    def source(self, a, U, x):
        # artery_network.py:303-326
        rpi = np.sqrt(np.pi)
        S2 = (-2 * rpi / a.db / a.param['Re'] * U[1] / np.sqrt(U[0])
              + (2 * np.sqrt(U[0]) * (rpi * a.f(x) + np.sqrt(a.A0(x)) * a.dfdr(x))
                 - U[0] * a.dfdr(x)) * a.drdx(x))
        return np.array([0, S2])

    def compute_U_half(self, a, x0, x1, U0, U1):
        # artery_network.py:329-358 -- dt/(x1-x0), dt/4
        dF = self.flux(a, U1, x1) - self.flux(a, U0, x0)
        sS = self.source(a, U0, x0) + self.source(a, U1, x1)
        return (U0 + U1) / 2 - a.dt / (x1 - x0) * dF + a.dt / 4 * sS

    def windkessel(self, a, k_max=100, tol=1.0e-12):
        # artery_network.py:361-422
        L = a.param['L']
        UL = a.Un(L)
        a.adjust_dex(L, UL[0], UL[1])
        xs = (L - 2 * a.dex, L - a.dex, L)
        Um2, Um1, Um0 = a.Un(xs[0]), a.Un(xs[1]), a.Un(xs[2])
        Uh21 = self.compute_U_half(a, xs[0], xs[1], Um2, Um1)
        Uh10 = self.compute_U_half(a, xs[1], xs[2], Um1, Um0)
        x21, x10 = L - 1.5 * a.dex, L - 0.5 * a.dex
        Fh21, Sh21 = self.flux(a, Uh21, x21), self.source(a, Uh21, x21)
        Fh10, Sh10 = self.flux(a, Uh10, x10), self.source(a, Uh10, x10)
        qm1 = Um1[1] - a.dt / a.dex * (Fh10[1] - Fh21[1]) + a.dt / 2 * (Sh10[1] + Sh21[1])
        pn = a.compute_outlet_pressure(Um0[0])
        p = pn
        R1 = a.param['R1'] * self.wk_scale[0]
        R2 = a.param['R2'] * self.wk_scale[1]
        CT = a.param['CT']
        its = 0
        for k in range(k_max):
            p_old = p
            qm0 = (Um0[1] + (p - pn) / R1 + self.dt / R1 / R2 / CT * pn
                   - self.dt * (R1 + R2) / R1 / R2 / CT * Um0[1])
            Am0 = Um0[0] - self.dt / a.dex * (qm0 - qm1)
            p = a.compute_outlet_pressure(Am0)
            its += 1
            if abs(p - p_old) < tol:
                break
        self._last_picard_its = its
        return Am0

    compute_A_out = windkessel        # pre-HEAD name

    def initial_x(self, p, d1, d2):
        # artery_network.py:434-461
        x = np.zeros(18)
        for v, (a, xe) in enumerate(((p, p.param['L']), (d1, 0), (d2, 0))):
            x[3 * v:3 * v + 3] = a.q0
            x[9 + 3 * v:12 + 3 * v] = a.A0(xe)
        return x

    def define_x(self):
        # artery_network.py:464-473
        for k, ip in enumerate(self.range_parent_arteries):
            i1, i2 = self.daughter_arteries(ip)
            self.x[k] = self.initial_x(self.arteries[ip], self.arteries[i1], self.arteries[i2])

    def _junction_ends(self, p, d1, d2):
        """(vessel, end point, interior point, ghost half point, interior half
        point, Jacobian ghost point, sign) for parent outlet / daughter inlets:
        artery_network.py:503-524 and :589-597."""
        pL = p.param['L']
        return ((p, pL, pL - p.dex, pL + p.dex / 2, pL - p.dex / 2, pL + p.dex, 1.0),
                (d1, 0, d1.dex, -d1.dex / 2, d1.dex / 2, -d1.dex, -1.0),
                (d2, 0, d2.dex, -d2.dex / 2, d2.dex / 2, -d2.dex, -1.0))

    def problem_function(self, p, d1, d2, x):
        # artery_network.py:476-562: the 18 residuals of a bifurcation
        y = np.zeros(18)
        press = []
        for v, (a, xe, xi, xg, xh, _, s) in enumerate(self._junction_ends(p, d1, d2)):
            Ue, Ui = a.Un(xe), a.Un(xi)
            if s > 0:
                Uh = self.compute_U_half(a, xi, xe, Ui, Ue)
            else:
                Uh = self.compute_U_half(a, xe, xi, Ue, Ui)
            ghost = np.array([x[11 + 3 * v], x[2 + 3 * v]])
            Fg, Sg = self.flux(a, ghost, xg), self.source(a, ghost, xg)
            Fh, Sh = self.flux(a, Uh, xh), self.source(a, Uh, xh)
            y[v] = 2 * x[3 * v + 1] - Uh[1] - x[3 * v + 2]                      # eq (20)
            y[3 + v] = 2 * x[10 + 3 * v] - Uh[0] - x[11 + 3 * v]                # eq (21)
            y[12 + v] = (x[3 * v] - Ue[1] + s * a.dt / a.dex * (Fg[1] - Fh[1])
                         - a.dt / 2 * (Sg[1] + Sh[1]))                          # eq (26)
            y[15 + v] = x[9 + 3 * v] - Ue[0] + s * a.dt / a.dex * (Fg[0] - Fh[0])  # eq (27)
            fe, A0e = a.f(xe), a.A0(xe)
            press.append((fe * (1 - np.sqrt(A0e / x[10 + 3 * v])),
                          fe * (1 - np.sqrt(A0e / x[9 + 3 * v]))))
        y[6] = x[0] - x[3] - x[6]                                               # eq (22)
        y[7] = x[1] - x[4] - x[7]
        y[8] = press[0][0] - press[1][0]                                        # eq (23), no p0
        y[9] = press[0][0] - press[2][0]
        y[10] = press[0][1] - press[1][1]
        y[11] = press[0][1] - press[2][1]
        return y

    def jacobian(self, p, d1, d2, x):
        # artery_network.py:565-665: hand-coded, 41 non-zeros, quirks verbatim
        J = np.zeros([18, 18])
        rpi = np.sqrt(np.pi)
        dt_dex_d1 = d1.dt / d1.dex
        for v, (a, xe, _, _, _, xj, s) in enumerate(self._junction_ends(p, d1, d2)):
            iq, iqh, iqg = 3 * v, 3 * v + 1, 3 * v + 2
            iA, iAh, iAg = 9 + 3 * v, 10 + 3 * v, 11 + 3 * v
            J[v, iqh], J[v, iqg] = 2, -1
            J[3 + v, iAh], J[3 + v, iAg] = 2, -1
            fe, A0e = a.f(xe), a.A0(xe)
            fh, A0h, dfdr, drdx = a.f(xj), a.A0(xj), a.dfdr(xj), a.drdx(xj)
            sgn = 1.0 if v == 0 else -1.0
            dP_half = sgn * fe * np.sqrt(A0e) / 2 / x[iAh] ** (3 / 2)
            dP_full = sgn * fe * np.sqrt(A0e) / 2 / x[iA] ** (3 / 2)
            if v == 0:
                J[8, iAh] = J[9, iAh] = dP_half
                J[10, iA] = J[11, iA] = dP_full
            else:
                J[7 + v, iAh] = dP_half
                J[9 + v, iA] = dP_full
            Re = a.param['Re']
            dtdex = a.dt / a.dex
            qg, Ag = x[iqg], x[iAg]
            J[12 + v, iq] = 1
            J[12 + v, iqg] = s * dtdex * 2 * qg / Ag + a.dt * rpi / a.db / Re / np.sqrt(Ag)
            J[12 + v, iAg] = (s * dtdex * (-(qg / Ag) ** 2 + fh / 2 * np.sqrt(A0h / Ag))
                              - a.dt / 2 * (rpi / a.db / Re * qg / Ag ** (3 / 2)
                                            + (1 / np.sqrt(Ag) * (rpi * fh + np.sqrt(A0h) * dfdr)
                                               - dfdr) * drdx))
            # row 17 uses d1's dt/dex (artery_network.py:662)
            J[15 + v, iqg] = s * (dtdex if v < 2 else dt_dex_d1)
            J[15 + v, iA] = 1
        J[6, 0], J[6, 3], J[6, 6] = 1, -1, -1
        J[7, 1], J[7, 4], J[7, 7] = 1, -1, -1
        return J

    def newton(self, p, d1, d2, x=None, k_max=30, tol=1.e-10):
        # artery_network.py:668-710
        x = np.ones(18) if x is None else x
        solves = 0
        for k in range(k_max):
            J = self.jacobian(p, d1, d2, x)
            func = self.problem_function(p, d1, d2, x)
            if np.linalg.norm(func) < tol:
                break
            try:
                x -= np.linalg.solve(J, func)
            except np.linalg.LinAlgError:
                eps = 1.e-6
                J += eps * np.eye(18)
                func[0] += eps
                x -= np.linalg.solve(J, func)
            solves += 1
        self._last_junction_solves = solves
        return x

    def adjust_bifurcation_step(self, p, d1, d2, margin=0.05):
        # artery_network.py:713-737
        pL = p.param['L']
        Up, U1, U2 = p.Un(pL), d1.Un(0), d2.Un(0)
        M = min([p.CFL_term(pL, Up[0], Up[1]), d1.CFL_term(0, U1[0], U1[1]),
                 d2.CFL_term(0, U2[0], U2[1])])
        p.dex = d1.dex = d2.dex = (1 + margin) * self.dt / M

    def set_inner_bc(self, ip, i1, i2):
        # artery_network.py:740-764
        p, d1, d2 = self.arteries[ip], self.arteries[i1], self.arteries[i2]
        self.adjust_bifurcation_step(p, d1, d2)
        k = self.range_parent_arteries.index(ip)
        self.x[k] = self.newton(p, d1, d2, self.x[k])
        self.junction_solves[k] = self._last_junction_solves
        p.U_out = np.array([self.x[k, 9], self.x[k, 0]])
        d1.U_in = np.array([self.x[k, 12], self.x[k, 3]])
        d2.U_in = np.array([self.x[k, 15], self.x[k, 6]])

    def set_bcs(self, q_in):
        # artery_network.py:767-786
        self.arteries[0].q_in = q_in
        for ip in self.range_parent_arteries:
            i1, i2 = self.daughter_arteries(ip)
            self.set_inner_bc(ip, i1, i2)
        for j, i in enumerate(self.range_leaf_arteries):
            self.arteries[i].A_out = self.windkessel(self.arteries[i])
            self.picard_its[j] = self._last_picard_its

    def step(self, n):
        """One pass of the loop body artery_network.py:916-944 (no dump)."""
        Nt = self.geo['Nt']
        self.set_bcs(self.q_ins[(n + 1) % Nt])
        for a in self.arteries:
            if a is not None:
                a.solve()
                a.update_solution()

    def solve(self, n_steps=None, record_counters=False):
        """artery_network.py:857-946 without the file output: returns the
        stored snapshots {'t', 'flow', 'area', 'pressure'} as arrays
        [snapshot, artery slot, node] (NaN rows for missing arteries)."""
        self.define_x()
        Nt, N_cycles = self.geo['Nt'], self.geo['N_cycles']
        Nt_store, N_cycles_store = self.sol['Nt_store'], self.sol['N_cycles_store']
        Nx = self.geo['Nx']
        out = {'t': [], 'flow': [], 'area': [], 'pressure': []}
        counters = {'artery_its': [], 'junction_solves': [], 'picard_its': []}
        t = 0
        done = 0
        for n_cycle in range(N_cycles):
            for n in range(Nt):
                if n_steps is not None and done >= n_steps:
                    break
                cycle_store = (n_cycle >= N_cycles - N_cycles_store)
                self.set_bcs(self.q_ins[(n + 1) % Nt])
                if cycle_store and n % (Nt / Nt_store) == 0:
                    snap = np.full((3, self.N, Nx + 1), np.nan)
                    for i, a in enumerate(self.arteries):
                        if a is not None:
                            a.update_pressure()
                            snap[0, i] = a._Un[:, 1]
                            snap[1, i] = a._Un[:, 0]
                            snap[2, i] = a.pn_values
                    out['t'].append(t)
                    out['flow'].append(snap[0])
                    out['area'].append(snap[1])
                    out['pressure'].append(snap[2])
                for a in self.arteries:
                    if a is not None:
                        a.solve()
                        a.update_solution()
                if record_counters:
                    counters['artery_its'].append(
                        [a.newton_its if a is not None else -1 for a in self.arteries])
                    counters['junction_solves'].append(self.junction_solves.copy())
                    counters['picard_its'].append(self.picard_its.copy())
                t += self.dt
                done += 1
        res = {k: np.array(v) for k, v in out.items()}
        if record_counters:
            res.update({k: np.array(v) for k, v in counters.items()})
        return res


def network_from_cfg(path, wk_scale=(0.3, 0.01), **overrides):
    """Build an OracleNetwork from a reference-format .cfg file.  ``overrides``
    are applied to whichever section holds the key (new keys go to
    [Parameters])."""
    param, geo, sol = load_cfg(path)
    for k, v in overrides.items():
        if k in geo:
            geo[k] = v
        elif k in sol:
            sol[k] = v
        else:
            param[k] = v
    return OracleNetwork(param, geo, sol, wk_scale=wk_scale)
