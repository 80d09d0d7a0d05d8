"""Artery -- the reference's per-vessel class (arteryfe/artery.py:15-415) as a
facade over the device-resident state.

The object holds host-side parameters only.  ``Un``/``U`` are views of device
memory; ``solve``/``update_solution`` and every boundary quantity run in the
CUDA library.  An Artery built on its own (as the reference's unit tests do)
owns a private one-vessel device network; inside an ArteryNetwork it is
attached to the network's handle.
"""
import math

import numpy as np

from .. import _ffi
from ..engine import DeviceNetwork


class _StateView(object):
    """``a.Un`` / ``a.U``: callable P1 function of (A, q) living on the device
    ([3P] dolfin.Function with allow_extrapolation, ar:147-158)."""

    def __init__(self, artery, which):
        self._a, self._which = artery, which

    def __call__(self, x):
        a = self._a
        return a._dev.eval_state(a._k, [x], which=self._which, network=a._b)[0]

    def values(self):
        """Vertex values, shape (Nx+1, 2) = (A, q), ascending x."""
        a = self._a
        xs = a.dx * np.arange(a.Nx + 1)
        out = [a._dev.eval_state(a._k, xs[i:i + 1024], which=self._which, network=a._b)
               for i in range(0, len(xs), 1024)]
        return np.concatenate(out)

    def split(self, deepcopy=False):
        v = self.values()
        return v[:, 0].copy(), v[:, 1].copy()


class _NodalFunction(object):
    """``a.pn``: host copy of a P1 scalar (the pressure of the last update)."""

    def __init__(self, dx, values):
        self.dx, self._v = dx, np.array(values, dtype=float)

    def values(self):
        return self._v

    def __call__(self, x):
        n = len(self._v) - 1
        j = min(max(int(np.floor(x / self.dx)), 0), n - 1)
        xi = (x - j * self.dx) / self.dx
        return self._v[j] * (1.0 - xi) + self._v[j + 1] * xi


class Artery(object):
    """One vessel: geometry, elasticity Expressions, (A, q) state, Dirichlet
    holders, Newton solve of the theta-scheme weak form.  Same constructor and
    method names as the reference."""

    def __init__(self, i, T, param, root=False, leaf=False):
        self.i, self.T = i, T
        self.root, self.leaf = root, leaf
        # the reference deep-copies (ar:61); entries are only ever replaced, never mutated in place, so a
        # per-artery dict over shared arrays is equivalent and keeps construction of large trees cheap
        self.param = dict(param)
        for key in ('Ru', 'Rd', 'L'):
            self.param[key] = param[key][i]
        self._dev = None      # DeviceNetwork
        self._k = 0           # compact artery index on the device
        self._b = 0           # ensemble member
        self._own = False
        self._dex_host = None
        self._in_network = False

    # ---- Expressions (ar:110-118): closed forms, used for set-up --------------
    def r0(self, x):
        return self._Ru * np.power(self._Rd / self._Ru, x / self._L)

    def A0(self, x):
        return np.pi * np.power(self.r0(x), 2)

    def f(self, x):
        p = self.param
        return 4.0 / 3.0 * (p['k1'] * np.exp(p['k2'] * self.r0(x)) + p['k3'])

    def dfdr(self, x):
        p = self.param
        return 4.0 / 3.0 * p['k1'] * p['k2'] * np.exp(p['k2'] * self.r0(x))

    def drdx(self, x):
        return np.log(self._Rd / self._Ru) / self._L * self.r0(x)

    def define_geometry(self, geo):
        """ar:67-118."""
        p = self.param
        self.Nx, self.Nt = geo['Nx'], geo['Nt']
        self._L, self._Ru, self._Rd = p['L'], p['Ru'], p['Rd']
        self.dx = self._L / self.Nx
        self.dt = self.T / self.Nt
        self._dex_host = self.dx
        self.db = math.sqrt(p['nu'] * self.T / 2 / np.pi)      # ar:100 (IEEE sqrt, same bits as np.sqrt)
        if self.leaf and self._Rd == 1:
            self._Rd = p['R_term']

    def define_solution(self, q0, theta=0.5, bc_tol=1.e-14):
        """ar:121-189.  Stand-alone arteries get their own device network."""
        self.q0, self.theta = q0, theta
        self._pn = None                  # pn = p0 everywhere (ar:161-162), materialised on first use
        if (self._dev is None and not self._in_network) or self._own:
            self._create_private_device()

    def _create_private_device(self):
        p = self.param
        if self._dev is not None:
            self._dev.close()
        wk = [float(np.ravel(p.get(k, 1.0))[0]) for k in ('R1', 'R2', 'CT')]
        self._dev = DeviceNetwork(
            self.Nx, self.Nt, self.theta, self.dt, junctions=np.zeros((0, 3)), leaves=[0] if self.leaf else [],
            k1=p['k1'], k2=p['k2'], k3=p['k3'], rho=p['rho'], Re=p['Re'], db=self.db, p0=p['p0'],
            Ru=[[self._Ru]], Rd=[[self._Rd]], L=[[self._L]], q0=[[self.q0]],
            R1=wk[0] * 0.3, R2=wk[1] * 0.01, CT=wk[2], q_ins=np.full(self.Nt, self.q0),
            root_artery=0 if self.root else -1)
        self._k, self._b, self._own = 0, 0, True

    def _attach(self, dev, k, b=0):
        self._dev, self._k, self._b, self._own, self._in_network = dev, k, b, False, True

    # ---- state -------------------------------------------------------------------
    @property
    def pn(self):
        if self._pn is None:
            self._pn = _NodalFunction(self.dx, np.full(self.Nx + 1, self.param['p0']))
        return self._pn

    @pn.setter
    def pn(self, value):
        self._pn = value

    @property
    def Un(self):
        return _StateView(self, 0)

    @property
    def U(self):
        return _StateView(self, 1)

    @property
    def dex(self):
        if self._dev is None:
            return self._dex_host
        return float(self._dev.get_dex()[self._b, self._k])

    @dex.setter
    def dex(self, value):
        if self._dev is None:
            self._dex_host = value
            return
        d = self._dev.get_dex()
        d[self._b, self._k] = value
        self._dev.set_dex(d)

    def solve(self):
        """ar:233-244: Newton on the variational form, on the device.  Raises
        RuntimeError on non-convergence like DOLFIN's NewtonSolver."""
        self._dev.artery_solve(self._k, self._b)
        if int(self._dev.get_flags()[self._b]) & _ffi.AF_FLAG_ARTERY_NOCONV:
            self._dev.clear_flags()           # the bits are sticky; this failure has been reported
            raise RuntimeError("Newton solver did not converge for artery %d" % self.i)

    def update_solution(self):
        """ar:247-251."""
        self._dev.artery_update(self._k, self._b)

    def update_pressure(self):
        """ar:254-261."""
        self.pn = _NodalFunction(self.dx, self._dev.get_pressure()[self._b, self._k])

    def compute_pressure(self, f, A0, A):
        """ar:264-282 (the reference reads a non-existent self.p0 here)."""
        return self.param['p0'] + f * (1 - np.sqrt(A0 / A))

    def compute_outlet_pressure(self, A):
        """ar:285-301."""
        L = self.param['L']
        return self.param['p0'] + self.f(L) * (1 - np.sqrt(self.A0(L) / A))

    def CFL_term(self, x, A, q):
        """ar:304-325, evaluated by the device code."""
        return self._dev.eval_cfl_term(self._k, x, A, q, network=self._b)

    def check_CFL(self, x, A, q):
        """ar:328-346."""
        return self.dt / self.dex < self.CFL_term(x, A, q)

    def adjust_dex(self, x, A, q, margin=0.05):
        """ar:349-365."""
        self.dex = (1 + margin) * self.dt / self.CFL_term(x, A, q)

    # ---- Dirichlet holders (ar:368-415) ------------------------------------------
    def _bc(self, lo, hi):
        return self._dev.get_bc()[self._b, self._k, lo:hi]

    def _set_bc(self, lo, values):
        bc = self._dev.get_bc()
        values = np.atleast_1d(np.asarray(values, dtype=float))
        bc[self._b, self._k, lo:lo + len(values)] = values
        self._dev.set_bc(bc)

    @property
    def q_in(self):
        return float(self._bc(1, 2)[0])

    @q_in.setter
    def q_in(self, value):
        self._set_bc(1, value)

    @property
    def U_in(self):
        return self._bc(0, 2).copy()

    @U_in.setter
    def U_in(self, U):
        self._set_bc(0, U)

    @property
    def A_out(self):
        return float(self._bc(2, 3)[0])

    @A_out.setter
    def A_out(self, value):
        self._set_bc(2, value)

    @property
    def U_out(self):
        return self._bc(2, 4).copy()

    @U_out.setter
    def U_out(self, U):
        self._set_bc(2, U)
