"""DeviceNetwork: a thin numpy-facing wrapper of one ``af_net`` handle.

It owns nothing but the opaque handle; every method is one C-ABI call (see
include/arteryfe_b200.h for the reference lines each call replaces).  Shapes:
``nb`` networks (ensemble members) x ``na`` arteries x ``Nx+1`` vertices.
"""
import ctypes

import numpy as np

from . import _ffi
from ._ffi import check, dptr, f64, i32, iptr, lib


class _PinnedResult(object):
    """Owner of a pinned result buffer of the library; numpy arrays made from it keep it alive and the
    buffer goes back to the library's host cache when the last of them is gone."""

    def __init__(self, ptr, shape):
        self._ptr = ptr
        self.__array_interface__ = {'shape': tuple(int(s) for s in shape), 'typestr': '<f8',
                                    'data': (ptr, False), 'version': 3}

    def __del__(self):
        try:
            lib.af_host_release(self._ptr)
        except Exception:
            pass


class DeviceNetwork(object):
    PINNED_STORE_BYTES = 64 << 20        # stores up to this size get library-pinned result arrays

    def __init__(self, Nx, Nt, theta, dt, junctions, leaves, k1, k2, k3, rho, Re, db, p0, Ru, Rd, L, q0,
                 R1, R2, CT, q_ins, device=0, stream=None, root_artery=0, **constants):
        Ru = np.atleast_2d(f64(Ru))
        self.nb, self.na = Ru.shape
        junctions = i32(junctions).reshape(-1, 3)
        leaves = i32(leaves).reshape(-1)
        self.nj, self.nl = len(junctions), len(leaves)
        self.Nx, self.Nt = int(Nx), int(Nt)
        self.np = self.Nx + 1

        def per_net(v):
            return f64(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.nb,)))

        def per_art(v):
            return f64(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.nb, self.na)))

        def per_leaf(v):
            return f64(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.nb, self.nl)))

        q_ins = f64(q_ins)
        if q_ins.ndim == 2 and q_ins.shape[0] != self.nb:
            raise ValueError("q_ins must be [Nt] or [n_networks, Nt]")
        if q_ins.shape[-1] != self.Nt:
            raise ValueError("q_ins must have Nt entries")
        # keep every array alive until af_net_create has copied it
        keep = dict(
            jp=i32(junctions[:, 0]), j1=i32(junctions[:, 1]), j2=i32(junctions[:, 2]), lf=leaves,
            k1=per_net(k1), k2=per_net(k2), k3=per_net(k3), rho=per_net(rho), Re=per_net(Re), db=per_net(db),
            p0=per_net(p0), Ru=Ru, Rd=per_art(Rd), L=per_art(L), q0=per_art(q0),
            R1=per_leaf(R1), R2=per_leaf(R2), CT=per_leaf(CT), q_ins=q_ins)
        d = _ffi.AfDesc()
        lib.af_desc_defaults(ctypes.byref(d))
        d.n_networks, d.n_arteries, d.n_junctions, d.n_leaves = self.nb, self.na, self.nj, self.nl
        d.Nx, d.Nt, d.theta, d.dt = self.Nx, self.Nt, float(theta), float(dt)
        d.q_ins_per_network = 1 if q_ins.ndim == 2 else 0
        d.junction_parent, d.junction_d1, d.junction_d2 = iptr(keep['jp']), iptr(keep['j1']), iptr(keep['j2'])
        d.leaf_artery = iptr(keep['lf'])
        for name in ('k1', 'k2', 'k3', 'rho', 'Re', 'db', 'p0', 'Ru', 'Rd', 'L', 'q0', 'R1', 'R2', 'CT', 'q_ins'):
            setattr(d, name, dptr(keep[name]))
        for name, value in constants.items():
            if not hasattr(d, name):
                raise TypeError("unknown solver constant %r" % name)
            setattr(d, name, value)
        d.device = int(device)
        d.root_artery = int(root_artery)
        d.stream = stream
        self._h = _ffi.Handle()
        check(lib.af_net_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.device = int(device)
        self._store_fields = 0

    # -- life cycle -------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None):
            lib.af_net_destroy(self._h)
            self._h = None
        self._host_store = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- time loop ----------------------------------------------------------------
    def step(self, n_first, n_steps):
        check(lib.af_net_step(self._h, int(n_first), int(n_steps)))

    def sync(self):
        check(lib.af_net_sync(self._h))

    def failed_step(self):
        """Global step at which an artery Newton solve first failed, or -1 (waits for the stream)."""
        s = ctypes.c_int64(-1)
        check(lib.af_net_failed_step(self._h, ctypes.byref(s)))
        return s.value

    def clear_flags(self):
        check(lib.af_net_clear_flags(self._h))

    def schedule(self):
        """How af_net_step runs the loop on this handle (_ffi.AF_SCHEDULE_*)."""
        k = ctypes.c_int32(0)
        check(lib.af_net_schedule(self._h, ctypes.byref(k)))
        return k.value

    def outlet_moments_device(self):
        """(device pointer, n_snapshots) of the packed moments [5][n_snapshots*n_leaves] of the stored
        outlet pressures over this handle's members (count, sum, sum of squares, min, max)."""
        p, n = ctypes.c_void_p(), ctypes.c_int32(0)
        check(lib.af_net_outlet_moments_d(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def set_store(self, step_mask, first_cycle, max_snapshots, fields):
        mask = np.ascontiguousarray(step_mask, dtype=np.uint8)
        assert mask.shape == (self.Nt,)
        check(lib.af_net_set_store(self._h, mask.ctypes.data_as(_ffi.P_u8), int(first_cycle), int(max_snapshots),
                                   int(fields)))
        self._store_fields = int(fields)
        self._host_store = None
        self._max_snap = int(max_snapshots)

    def stream_store_to_host(self):
        """Allocate the host arrays of the configured store now and let af_net_step
        stream every snapshot into them while the loop runs (af_net_set_store_host)."""
        n = self._max_snap
        out, ptrs = {}, []
        layout = (('flow', _ffi.AF_STORE_FLOW, (n, self.nb, self.na, self.np)),
                  ('area', _ffi.AF_STORE_AREA, (n, self.nb, self.na, self.np)),
                  ('pressure', _ffi.AF_STORE_PRESSURE, (n, self.nb, self.na, self.np)),
                  ('outlet_pressure', _ffi.AF_STORE_OUTLET_PRESSURE, (n, self.nb, self.nl)))
        total = sum(8 * int(np.prod(shape)) for _, bit, shape in layout if self._store_fields & bit)
        if total <= self.PINNED_STORE_BYTES:
            # small store: result arrays in pinned memory of the library, snapshots land in them directly
            raw = [_ffi.P_dbl() for _ in layout]
            check(lib.af_net_set_store_host_pinned(self._h, *[ctypes.byref(r) for r in raw]))
            for (name, bit, shape), r in zip(layout, raw):
                if self._store_fields & bit:
                    out[name] = np.asarray(_PinnedResult(ctypes.cast(r, ctypes.c_void_p).value, shape))
            self._host_store = out
            return
        for name, bit, shape in (('flow', _ffi.AF_STORE_FLOW, (n, self.nb, self.na, self.np)),
                                 ('area', _ffi.AF_STORE_AREA, (n, self.nb, self.na, self.np)),
                                 ('pressure', _ffi.AF_STORE_PRESSURE, (n, self.nb, self.na, self.np)),
                                 ('outlet_pressure', _ffi.AF_STORE_OUTLET_PRESSURE, (n, self.nb, self.nl))):
            if self._store_fields & bit:
                out[name] = np.empty(shape)
                ptrs.append(dptr(out[name]))
            else:
                ptrs.append(None)
        check(lib.af_net_set_store_host(self._h, *ptrs))
        self._host_store = out

    def snapshot_count(self):
        c = ctypes.c_int32(0)
        check(lib.af_net_snapshot_count(self._h, ctypes.byref(c)))
        return c.value

    def fetch_snapshots(self):
        n = self.snapshot_count()
        if getattr(self, '_host_store', None) is not None:
            hs = self._host_store
            step = np.zeros(n, dtype=np.int64)
            check(lib.af_net_fetch_snapshots(
                self._h, *[dptr(hs[k]) if k in hs else None for k in ('flow', 'area', 'pressure', 'outlet_pressure')],
                step.ctypes.data_as(_ffi.P_i64)))
            out = {k: v[:n] for k, v in hs.items()}
            out['step'] = step
            return out
        out = {'step': np.zeros(n, dtype=np.int64)}
        ptrs = []
        for name, bit in (('flow', _ffi.AF_STORE_FLOW), ('area', _ffi.AF_STORE_AREA),
                          ('pressure', _ffi.AF_STORE_PRESSURE)):
            if self._store_fields & bit:
                out[name] = np.zeros((n, self.nb, self.na, self.np))
                ptrs.append(dptr(out[name]))
            else:
                ptrs.append(None)
        if self._store_fields & _ffi.AF_STORE_OUTLET_PRESSURE:
            out['outlet_pressure'] = np.zeros((n, self.nb, self.nl))
            ptrs.append(dptr(out['outlet_pressure']))
        else:
            ptrs.append(None)
        check(lib.af_net_fetch_snapshots(self._h, ptrs[0], ptrs[1], ptrs[2], ptrs[3],
                                         out['step'].ctypes.data_as(_ffi.P_i64)))
        return out

    def outlet_pressure_device_ptr(self):
        p = ctypes.c_void_p()
        check(lib.af_net_outlet_pressure_d(self._h, ctypes.byref(p)))
        return p.value

    # -- state ---------------------------------------------------------------------
    def get_state(self):
        A = np.zeros((self.nb, self.na, self.np))
        q = np.zeros_like(A)
        check(lib.af_net_get_state(self._h, dptr(A), dptr(q)))
        return A, q

    def set_state(self, A, q):
        A, q = f64(A).reshape(self.nb, self.na, self.np), f64(q).reshape(self.nb, self.na, self.np)
        check(lib.af_net_set_state(self._h, dptr(A), dptr(q)))

    def get_pressure(self):
        p = np.zeros((self.nb, self.na, self.np))
        check(lib.af_net_get_pressure(self._h, dptr(p)))
        return p

    def _get(self, fn, shape, dtype=np.float64):
        a = np.zeros(shape, dtype=dtype)
        check(fn(self._h, a.ctypes.data_as(ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype)))))
        return a

    def get_bc(self):
        return self._get(lib.af_net_get_bc, (self.nb, self.na, 4))

    def set_bc(self, bc):
        bc = f64(bc).reshape(self.nb, self.na, 4)
        check(lib.af_net_set_bc(self._h, dptr(bc)))

    def get_dex(self):
        return self._get(lib.af_net_get_dex, (self.nb, self.na))

    def set_dex(self, dex):
        dex = f64(dex).reshape(self.nb, self.na)
        check(lib.af_net_set_dex(self._h, dptr(dex)))

    def get_junction_x(self):
        return self._get(lib.af_net_get_junction_x, (self.nb, self.nj, 18))

    def set_junction_x(self, x):
        x = f64(x).reshape(self.nb, self.nj, 18)
        check(lib.af_net_set_junction_x(self._h, dptr(x)))

    def get_counters(self):
        a = np.zeros((self.nb, self.na), dtype=np.int32)
        j = np.zeros((self.nb, self.nj), dtype=np.int32)
        p = np.zeros((self.nb, self.nl), dtype=np.int32)
        check(lib.af_net_get_counters(self._h, iptr(a), iptr(j), iptr(p)))
        return a, j, p

    def get_totals(self):
        t = np.zeros(3, dtype=np.int64)
        check(lib.af_net_get_totals(self._h, t.ctypes.data_as(_ffi.P_i64)))
        return t

    def get_flags(self):
        return self._get(lib.af_net_get_flags, (self.nb,), np.uint32)

    def set_counter_log(self, capacity):
        check(lib.af_net_set_counter_log(self._h, int(capacity)))
        self._log_cap = int(capacity)

    def fetch_counter_log(self):
        n = ctypes.c_int32(0)
        cap = self._log_cap
        a = np.zeros((cap, self.nb, self.na), dtype=np.int32)
        j = np.zeros((cap, self.nb, self.nj), dtype=np.int32)
        p = np.zeros((cap, self.nb, self.nl), dtype=np.int32)
        check(lib.af_net_fetch_counter_log(self._h, ctypes.byref(n), iptr(a), iptr(j), iptr(p)))
        return a[:n.value], j[:n.value], p[:n.value]

    # -- single phases / fine-grained evaluations -------------------------------
    def phase_boundary(self):
        check(lib.af_phase_boundary(self._h))

    def phase_arteries(self, inlet_index=-1):
        check(lib.af_phase_arteries(self._h, int(inlet_index)))

    def artery_solve(self, artery, network=0):
        check(lib.af_artery_solve(self._h, int(network), int(artery)))

    def artery_update(self, artery, network=0):
        check(lib.af_artery_update(self._h, int(network), int(artery)))

    def eval_state(self, artery, x, which=0, network=0):
        x = f64(np.atleast_1d(x))
        out = np.zeros((len(x), 2))
        check(lib.af_eval_state(self._h, int(network), int(artery), int(which), len(x), dptr(x), dptr(out)))
        return out

    def eval_coefficients(self, artery, x, network=0):
        x = f64(np.atleast_1d(x))
        out = np.zeros((len(x), 4))
        check(lib.af_eval_coefficients(self._h, int(network), int(artery), len(x), dptr(x), dptr(out)))
        return out

    def eval_flux_source(self, artery, x, U, network=0):
        x = f64(np.atleast_1d(x))
        U = f64(U).reshape(len(x), 2)
        out = np.zeros((len(x), 3))
        check(lib.af_eval_flux_source(self._h, int(network), int(artery), len(x), dptr(x), dptr(U), dptr(out)))
        return out

    def eval_u_half(self, artery, x0, x1, U0, U1, network=0):
        U0, U1, out = f64(U0), f64(U1), np.zeros(2)
        check(lib.af_eval_u_half(self._h, int(network), int(artery), float(x0), float(x1), dptr(U0), dptr(U1),
                                 dptr(out)))
        return out

    def eval_cfl_term(self, artery, x, A, q, network=0):
        M = ctypes.c_double(0)
        check(lib.af_eval_cfl_term(self._h, int(network), int(artery), float(x), float(A), float(q),
                                   ctypes.byref(M)))
        return M.value

    def junction_cfl_dex(self, junction, network=0):
        d = ctypes.c_double(0)
        check(lib.af_junction_cfl_dex(self._h, int(network), int(junction), ctypes.byref(d)))
        return d.value

    def junction_cfl_min(self, junction, network=0):
        m = ctypes.c_double(0)
        check(lib.af_junction_cfl_min(self._h, int(network), int(junction), ctypes.byref(m)))
        return m.value

    def junction_eval(self, junction, dex3, x, network=0):
        dex3, x = f64(np.broadcast_to(dex3, (3,))), f64(x)
        y, J = np.zeros(18), np.zeros((18, 18))
        check(lib.af_junction_eval(self._h, int(network), int(junction), dptr(dex3), dptr(x), dptr(y), dptr(J)))
        return y, J

    def junction_newton(self, junction, dex3, x, k_max=30, tol=1e-10, network=0):
        dex3, x = f64(np.broadcast_to(dex3, (3,))), f64(x).copy()
        s = ctypes.c_int32(0)
        check(lib.af_junction_newton(self._h, int(network), int(junction), dptr(dex3), dptr(x), int(k_max),
                                     float(tol), ctypes.byref(s)))
        return x, s.value

    def windkessel(self, leaf, k_max=100, tol=1e-12, network=0):
        A, d, its = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int32(0)
        check(lib.af_windkessel(self._h, int(network), int(leaf), int(k_max), float(tol), ctypes.byref(A),
                                ctypes.byref(d), ctypes.byref(its)))
        return A.value, d.value, its.value
