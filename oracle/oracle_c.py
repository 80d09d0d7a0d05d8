"""ctypes wrapper of oracle/oracle.c (TEST INFRASTRUCTURE ONLY, see
oracle/__init__.py).  The network description comes from the numpy oracle's
set-up (OracleNetwork), so both oracles start from identical inputs."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'liboracle_c.so')
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build():
    src = os.path.join(HERE, 'oracle.c')
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_create.restype = ctypes.c_void_p
        _lib.orc_create.argtypes = ([ctypes.c_int] * 5 + [ctypes.c_double] * 2 + [_ip] * 4
                                    + [ctypes.c_double] * 7 + [_dp] * 8)
        _lib.orc_destroy.argtypes = [ctypes.c_void_p]
        _lib.orc_step.restype = ctypes.c_int
        _lib.orc_step.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, _ip, _ip]
        for name in ('orc_get_state', 'orc_set_state'):
            getattr(_lib, name).argtypes = [ctypes.c_void_p, _dp, _dp]
        _lib.orc_get_pending.argtypes = [ctypes.c_void_p, _dp, _dp]
        _lib.orc_artery_assemble.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp, _dp]
        _lib.orc_artery_solve.restype = ctypes.c_int
        _lib.orc_artery_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        for name in ('orc_get_x', 'orc_get_bc', 'orc_get_pressure', 'orc_set_bc'):
            getattr(_lib, name).argtypes = [ctypes.c_void_p, _dp]
        _lib.orc_get_totals.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.intc)


class COracle(object):
    """C oracle of one network, built from an ``OracleNetwork`` ``net``
    (only its set-up is used: topology, nondimensional parameters, q0, q_ins)."""

    def __init__(self, net):
        L = lib()
        arts = [a for a in net.arteries if a is not None]
        compact = {a.i: k for k, a in enumerate(arts)}
        self.na, self.np = len(arts), arts[0].Nx + 1
        jn = [[compact[ip]] + [compact[i] for i in net.daughter_arteries(ip)] for ip in net.range_parent_arteries]
        jn = np.array(jn, dtype=np.intc).reshape(-1, 3)
        leaves = [compact[i] for i in net.range_leaf_arteries]
        self.nj, self.nl = len(jn), len(leaves)
        nd = net.nondim
        args = dict(
            jp=_i(jn[:, 0]), j1=_i(jn[:, 1]), j2=_i(jn[:, 2]), leaf=_i(leaves),
            Ru=_d([a.Ru for a in arts]), Rd=_d([a.Rd for a in arts]), L=_d([a.L for a in arts]),
            q0=_d([a.q0 for a in arts]),
            R1=_d([net.arteries[i].param['R1'] * net.wk_scale[0] for i in net.range_leaf_arteries]),
            R2=_d([net.arteries[i].param['R2'] * net.wk_scale[1] for i in net.range_leaf_arteries]),
            CT=_d([net.arteries[i].param['CT'] for i in net.range_leaf_arteries]),
            q_ins=_d(net.q_ins))
        self._keep = args
        ip = lambda a: a.ctypes.data_as(_ip)      # noqa: E731
        dp = lambda a: a.ctypes.data_as(_dp)      # noqa: E731
        self._h = L.orc_create(
            self.na, self.nj, self.nl, arts[0].Nx, net.geo['Nt'], net.dt, net.sol['theta'],
            ip(args['jp']), ip(args['j1']), ip(args['j2']), ip(args['leaf']),
            nd['k1'], nd['k2'], nd['k3'], nd['rho'], nd['Re'], arts[0].db, nd['p0'],
            dp(args['Ru']), dp(args['Rd']), dp(args['L']), dp(args['q0']), dp(args['R1']), dp(args['R2']),
            dp(args['CT']), dp(args['q_ins']))
        self.max_threads = L.orc_max_threads()

    def step(self, g0, count, nthreads=0, log=False):
        la = np.zeros((count, self.na), dtype=np.intc) if log else None
        lj = np.zeros((count, self.nj), dtype=np.intc) if log else None
        failed = lib().orc_step(self._h, g0, count, nthreads,
                                la.ctypes.data_as(_ip) if log else None, lj.ctypes.data_as(_ip) if log else None)
        return (failed, la, lj) if log else failed

    def state(self):
        A, q = np.zeros((self.na, self.np)), np.zeros((self.na, self.np))
        lib().orc_get_state(self._h, A.ctypes.data_as(_dp), q.ctypes.data_as(_dp))
        return A, q

    def set_state(self, A, q):
        A, q = _d(A), _d(q)
        lib().orc_set_state(self._h, A.ctypes.data_as(_dp), q.ctypes.data_as(_dp))

    def pending(self):
        A, q = np.zeros((self.na, self.np)), np.zeros((self.na, self.np))
        lib().orc_get_pending(self._h, A.ctypes.data_as(_dp), q.ctypes.data_as(_dp))
        return A, q

    def set_bc(self, bc):
        bc = _d(bc)
        assert bc.shape == (self.na, 4)
        lib().orc_set_bc(self._h, bc.ctypes.data_as(_dp))

    def bc(self):
        bc = np.zeros((self.na, 4))
        lib().orc_get_bc(self._h, bc.ctypes.data_as(_dp))
        return bc

    def artery_assemble(self, a, UA, Uq):
        """Residual (np, 2) and Jacobian blocks (3, np, 2, 2), Dirichlet rows applied."""
        UA, Uq = _d(UA), _d(Uq)
        R, J = np.zeros((self.np, 2)), np.zeros((3, self.np, 2, 2))
        lib().orc_artery_assemble(self._h, a, UA.ctypes.data_as(_dp), Uq.ctypes.data_as(_dp),
                                  R.ctypes.data_as(_dp), J.ctypes.data_as(_dp))
        return R, J

    def artery_solve(self, a, max_it=1000):
        return lib().orc_artery_solve(self._h, a, max_it)

    def x(self):
        x = np.zeros((self.nj, 18))
        lib().orc_get_x(self._h, x.ctypes.data_as(_dp))
        return x

    def pressure(self):
        p = np.zeros((self.na, self.np))
        lib().orc_get_pressure(self._h, p.ctypes.data_as(_dp))
        return p

    def totals(self):
        t = (ctypes.c_longlong * 3)()
        lib().orc_get_totals(self._h, t)
        return list(t)

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
