"""ArteryNetwork -- the reference's network driver
(arteryfe/artery_network.py:14-946) as a facade over one device handle.

Construction does the reference's host set-up (tree geometry, nondimensional
parameters, inlet table, topology lists, initial flow split) and uploads it
once.  ``solve()`` then runs the whole time loop device-resident through
``af_net_step``; the fine-grained methods (``flux``, ``problem_function``,
``newton``, ``windkessel`` ...) keep the reference's signatures and evaluate the
same device code the time loop uses, one call at a time.
"""
import configparser
import os
import weakref

import copy

import numpy as np

from .. import _ffi
from ..engine import DeviceNetwork
from .artery import Artery
from .utils import nondimensionalise_parameters, read_inlet, _SeriesFile

HEAD_WK_SCALE = (0.3, 0.01)          # artery_network.py:408-409


class _JunctionX(np.ndarray):
    """``an.x``: host copy of the junction unknowns that writes item
    assignments through to the device (the reference mutates ``self.x[i]``)."""

    def __new__(cls, values, dev):
        obj = np.asarray(values, dtype=float).view(cls)
        obj._dev = dev
        return obj

    def __array_finalize__(self, obj):
        self._dev = getattr(obj, '_dev', None)

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        if self._dev is not None and self.ndim == 2:
            self._dev.set_junction_x(np.asarray(self)[None])


class _Arteries(object):
    """``an.arteries``: the reference's heap-indexed list of Artery objects (``None`` where the tree
    has no vessel, an:87-93).  The facades carry no state of their own -- everything lives on the
    device -- so they are built on first access: constructing a 255-vessel network does not pay for
    255 Python objects it may never touch."""

    def __init__(self, net):
        # a weak reference: the network must die by reference count (no cycle), so that its device and
        # pinned blocks return to the library's caches as soon as the caller drops it
        self._net, self._made = weakref.proxy(net), {}

    def __len__(self):
        return self._net.N

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self._net.N))]
        if i < 0:
            i += self._net.N
        if not 0 <= i < self._net.N:
            raise IndexError(i)
        if not self._net._exists[i]:
            return None
        a = self._made.get(i)
        if a is None:
            a = self._made[i] = self._net._make_artery(i)
        return a

    def __iter__(self):
        return (self[i] for i in range(self._net.N))


class ArteryNetwork(object):

    def __init__(self, param, wk_scale=None, device=0, stream=None, _host_only=False):
        """artery_network.py:56-130.  ``param`` is a ParamParser (or anything
        with ``.param``, ``.geo``, ``.solution`` dicts).  ``wk_scale`` overrides
        the two hard-coded Windkessel factors of HEAD (0.3, 0.01); it can also
        be given as ``R1_factor`` / ``R2_factor`` in the [Solution] section."""
        # the reference overwrites param.param['Ru'/'Rd'/'L'/'R1'/...] in place (an:66-70), which
        # breaks a second ArteryNetwork built from the same ParamParser; work on a copy instead
        param = copy.copy(param)
        param.param = dict(param.param)
        order = param.param['order']
        self.N = 2 ** order - 1
        self.device = device
        self._stream = stream        # optional cudaStream_t (int) the device work is enqueued on
        self._host_only = _host_only  # set-up only (ensemble descriptions, CPU tests); no device handle

        if 'alpha' in param.param.keys():
            if 'R_term' not in param.param:
                raise KeyError("R_term: a geometry built from 'alpha' needs a terminal radius "
                               "(artery_network.py:66)")
            Ru, Rd, L = self.build_geometry(order, param.param['Ru'], param.param['Rd'],
                                            param.param['alpha'], param.param['L'], param.param['R_term'])
            param.param['Ru'], param.param['Rd'], param.param['L'] = Ru, Rd, L

        n_leaf_max = 2 ** (order - 1)
        for key in ('R1', 'R2', 'CT'):      # scalar in the shipped demo cfg => same value on every leaf
            if np.ndim(param.param[key]) == 0:
                param.param[key] = np.full(n_leaf_max, float(param.param[key]))

        self.nondim = nondimensionalise_parameters(param)
        self.geo = param.geo
        self.sol = param.solution
        if wk_scale is None:
            wk_scale = (self.sol.get('R1_factor', HEAD_WK_SCALE[0]), self.sol.get('R2_factor', HEAD_WK_SCALE[1]))
        self.wk_scale = tuple(wk_scale)

        rc, qc = self.nondim['rc'], self.nondim['qc']
        Nt = self.geo['Nt']
        self.T, self.q_ins = read_inlet(self.sol['inlet_flow_location'], Nt)
        self.T = self.T * qc / rc ** 3
        self.q_ins = self.q_ins / qc
        self.dt = self.T / Nt

        self.check_geometry()

        # an:87-118 -- which vessels exist, root / parent / daughter / leaf classification (the reference's
        # loop over the heap indices, for all vessels at once)
        self._exists = np.asarray(self.nondim['Ru']) != 0
        self.arteries = _Arteries(self)
        heap = np.arange(self.N)
        has = np.zeros((2, self.N), dtype=bool)
        for s in (0, 1):
            k = 2 * heap + 1 + s
            ok = k <= self.N - 1
            has[s, ok] = self._exists[k[ok]]
        self._has_daughter = has
        body = self._exists & (heap >= 1)
        self.range_daughter_arteries = heap[body].tolist()
        self.range_leaf_arteries = heap[body & ~has[0] & ~has[1]].tolist()
        self.range_parent_arteries = [0] + heap[body & (has[0] | has[1])].tolist()
        self._leaf_slot = {i: j for j, i in enumerate(self.range_leaf_arteries)}

        self._dev = None
        self.define_geometry()
        self.define_solution()

    # ---- topology (artery_network.py:133-192) -----------------------------------
    def daughter_arteries(self, i):
        out = []
        for k in (2 * i + 1, 2 * i + 2):
            out.append(k if k <= self.N - 1 and self._exists[k] else None)
        return out[0], out[1]

    def parent_artery(self, i):
        return (i - 1) // 2

    def sister_artery(self, i):
        return i + 1 if i % 2 else i - 1

    def build_geometry(self, order, Ru, Rd, alpha, L, R_term):
        """artery_network.py:195-217: radii shrink by ``alpha`` per generation
        (times the per-vessel multiplier in Ru/Rd), last-generation Rd is
        R_term, lengths are L*Ru, second daughters are 0.01 shorter."""
        Ru = np.array(Ru, dtype=float)
        Rd = np.array(Rd, dtype=float)
        lengths = np.zeros(self.N)
        lengths[0] = L * Ru[0]
        lo = hi = 0
        for level in range(order):
            last_generation = not (2 ** (level + 1) - 1 < len(Rd))
            # the reference's loop over the parents p of this generation, all of them at once (parents and
            # daughters of one generation are disjoint index ranges, so the order within it does not matter)
            p = np.arange(lo, hi)
            d1, d2 = 2 * p + 1, 2 * p + 2
            Ru[d1] = alpha * Ru[p] * Ru[d1]
            Ru[d2] = np.maximum(0, alpha * Ru[p]) * Ru[d2]
            if last_generation:
                Rd[d1] = Rd[d2] = R_term
            else:
                Rd[d1] = alpha * Rd[p] * Rd[d1]
                Rd[d2] = np.maximum(0, alpha * Rd[p]) * Rd[d2]
            lengths[d1] = L * Ru[d1]
            lengths[d2] = L * Ru[d2] - 0.01
            lo += int(2 ** (level - 1))
            hi += int(2 ** level)
        return Ru, Rd, lengths

    def check_geometry(self):
        """artery_network.py:220-241."""
        order = self.nondim['order']
        for key in ('Ru', 'Rd', 'L'):
            got = len(self.nondim[key])
            assert got == self.N, \
                "A network of order {} requires {} values for {}, {} were provided".format(order, self.N, key, got)
        n1, n2, nc = (len(np.atleast_1d(self.nondim[k])) for k in ('R1', 'R2', 'CT'))
        assert n1 == n2, "R2 must have the same number of values as R1. {} != {}".format(n2, n1)
        assert n1 == nc, "CT must have the same number of values as R1. {} != {}".format(nc, n1)

    # ---- set-up -------------------------------------------------------------------
    def _make_artery(self, i):
        """The facade of vessel ``i`` (an:87-93, 121-127 + Artery.define_geometry / define_solution)."""
        a = Artery(i, self.T, self.nondim, root=(i == 0), leaf=(i in self._leaf_slot))
        if a.leaf:
            j = self._leaf_slot[i]
            for key in ('R1', 'R2', 'CT'):
                a.param[key] = self.nondim[key][j]
        a._in_network = True
        a.define_geometry(self.geo)
        if getattr(self, '_q0', None) is not None:
            a._attach(self._dev, self._compact[i])
            a.define_solution(self._q0[i], self.sol['theta'])
        return a

    def define_geometry(self):
        """artery_network.py:244-255: per-vessel geometry (Artery.define_geometry, ar:67-118), for all
        vessels at once; facades built so far are refreshed."""
        nd = self.nondim
        idx = np.flatnonzero(self._exists)
        self._idx = idx
        Ru, Rd, L = (np.asarray(nd[k], dtype=float)[idx] for k in ('Ru', 'Rd', 'L'))
        is_leaf = np.isin(idx, self.range_leaf_arteries)
        if 'R_term' in nd:
            Rd = np.where(is_leaf & (Rd == 1), nd['R_term'], Rd)          # ar:108-109
        self._geom = dict(Ru=Ru, Rd=Rd, L=L)
        self._db = float(np.sqrt(nd['nu'] * self.T / 2 / np.pi))              # ar:100, the same for every vessel
        for a in self.arteries._made.values():
            a.define_geometry(self.geo)

    def define_solution(self):
        """artery_network.py:258-279: initial flows (root q_ins[0]; daughters
        split by inlet rest areas), then the device-resident state."""
        idx = self._idx
        self._compact = {i: k for k, i in enumerate(idx.tolist())}
        compact = np.full(self.N, -1, dtype=np.int64)
        compact[idx] = np.arange(len(idx))
        g = self._geom
        # Artery.A0(0) of every vessel: pi * (Ru * (Rd/Ru)**(0/L))**2, the facade's expressions (ar:110-112)
        A0_in = np.ones(self.N)
        A0_in[idx] = np.pi * np.power(g['Ru'] * np.power(g['Rd'] / g['Ru'], 0 / g['L']), 2)
        # an:262-276, one generation at a time (a daughter's share needs its parent's flow)
        q0_all = np.zeros(self.N)
        q0_all[0] = self.q_ins[0]
        heap = np.arange(self.N)
        sister = np.where(heap % 2 == 1, heap + 1, heap - 1)
        sister[0] = 0
        for level in range(1, self.nondim['order']):
            i = heap[2 ** level - 1:2 ** (level + 1) - 1]
            i = i[self._exists[i]]
            q0_all[i] = A0_in[i] / (A0_in[i] + A0_in[sister[i]]) * q0_all[(i - 1) // 2]
        q0 = dict(zip(idx.tolist(), q0_all[idx].tolist()))
        parents = np.asarray(self.range_parent_arteries, dtype=np.int64)
        has = self._has_daughter[:, parents]
        if not np.all(has[0] & has[1]):
            ip = int(parents[np.flatnonzero(~(has[0] & has[1]))[0]])
            raise ValueError("artery %d has exactly one daughter; a parent needs two "
                             "(artery_network.py:111-118)" % ip)
        junctions = np.stack([compact[parents], compact[2 * parents + 1], compact[2 * parents + 2]], axis=1)
        leaves = compact[np.asarray(self.range_leaf_arteries, dtype=np.int64)]
        self._q0_compact = q0_all[idx]
        self._q0 = q0
        self._description = self.describe(junctions, leaves, q0)
        if self._dev is not None:
            self._dev.close()
        self._dev = self._make_device(self._description)
        for a in self.arteries._made.values():
            a._attach(self._dev, self._compact[a.i])
            a.define_solution(q0[a.i], self.sol['theta'])
        if self._dev is not None:
            self.x = np.ones([len(self.range_parent_arteries), 18])       # an:279 (define_x replaces it, an:874)

    def describe(self, junctions, leaves, q0):
        """Flat, nondimensional description of the network: exactly the arrays
        af_desc takes (include/arteryfe_b200.h).  Also used by the ensemble
        driver, which stacks perturbed copies of it."""
        nd, g = self.nondim, self._geom
        R1, R2, CT = (np.asarray(nd[k], dtype=float)[:len(leaves)] for k in ('R1', 'R2', 'CT'))
        return dict(
            Nx=self.geo['Nx'], Nt=self.geo['Nt'], theta=self.sol['theta'], dt=self.dt,
            junctions=np.array(junctions, dtype=np.int32).reshape(-1, 3), leaves=np.array(leaves, dtype=np.int32),
            k1=nd['k1'], k2=nd['k2'], k3=nd['k3'], rho=nd['rho'], Re=nd['Re'], db=self._db, p0=nd['p0'],
            Ru=g['Ru'], Rd=g['Rd'], L=g['L'], q0=self._q0_compact.copy(),
            R1=R1 * self.wk_scale[0], R2=R2 * self.wk_scale[1], CT=CT,
            q_ins=self.q_ins)

    def _make_device(self, d):
        if self._host_only:
            return None
        return DeviceNetwork(
            d['Nx'], d['Nt'], d['theta'], d['dt'], d['junctions'], d['leaves'],
            k1=d['k1'], k2=d['k2'], k3=d['k3'], rho=d['rho'], Re=d['Re'], db=d['db'], p0=d['p0'],
            Ru=d['Ru'][None], Rd=d['Rd'][None], L=d['L'][None], q0=d['q0'][None],
            R1=d['R1'][None], R2=d['R2'][None], CT=d['CT'][None],
            q_ins=d['q_ins'], device=self.device, stream=self._stream)

    @property
    def x(self):
        return _JunctionX(self._dev.get_junction_x()[0], self._dev)

    @x.setter
    def x(self, value):
        self._dev.set_junction_x(np.asarray(value, dtype=float)[None])

    def _junction_of(self, p):
        return self.range_parent_arteries.index(p.i)

    def _leaf_of(self, a):
        return self.range_leaf_arteries.index(a.i)

    # ---- boundary physics, evaluated by the device code ---------------------------
    def flux(self, a, U, x):
        """artery_network.py:282-300."""
        return self._dev.eval_flux_source(a._k, [x], [U])[0, :2].copy()

    def source(self, a, U, x):
        """artery_network.py:303-326."""
        return np.array([0.0, self._dev.eval_flux_source(a._k, [x], [U])[0, 2]])

    def compute_U_half(self, a, x0, x1, U0, U1):
        """artery_network.py:329-358."""
        return self._dev.eval_u_half(a._k, x0, x1, U0, U1)

    def windkessel(self, a, k_max=100, tol=1.0e-12):
        """artery_network.py:361-422: outlet area of a leaf at t_(n+1); also
        leaves the CFL-scaled ``a.dex`` behind, as the reference does."""
        A_out, dex, its = self._dev.windkessel(self._leaf_of(a), k_max, tol)
        a.dex = dex
        self.last_picard_its = its
        return A_out

    compute_A_out = windkessel         # name used before HEAD (BASELINE.json, reference tests)

    def structured_tree(self, a):
        """artery_network.py:425-431 is an empty stub in the reference."""
        raise NotImplementedError("structured_tree is not implemented by the reference either")

    def initial_x(self, p, d1, d2):
        """artery_network.py:434-461."""
        x = np.zeros(18)
        ends = ((p, p.param['L']), (d1, 0), (d2, 0))
        for v, (a, xe) in enumerate(ends):
            x[3 * v:3 * v + 3] = a.q0
            x[9 + 3 * v:12 + 3 * v] = a.A0(xe)
        return x

    def define_x(self):
        """artery_network.py:464-473."""
        d = self._description
        jn = d['junctions']
        if len(jn) == 0:
            self.x = np.zeros((0, 18))
            return
        # initial_x for every junction at once: q0 x3 and A0(end) x3 per vessel (an:434-461)
        q0 = d['q0'][jn]                                           # [nj, 3]
        Ru, Rd = d['Ru'][jn], d['Rd'][jn]
        frac = np.zeros_like(Ru)
        frac[:, 0] = 1.0                                           # parent: x = L, daughters: x = 0
        A_end = np.pi * np.power(Ru * np.power(Rd / Ru, frac), 2)  # Artery.A0 at x/L = frac, same expressions
        x = np.concatenate([np.repeat(q0, 3, axis=1), np.repeat(A_end, 3, axis=1)], axis=1)
        self.x = x

    def problem_function(self, p, d1, d2, x):
        """artery_network.py:476-562."""
        return self._dev.junction_eval(self._junction_of(p), [p.dex, d1.dex, d2.dex], x)[0]

    def jacobian(self, p, d1, d2, x):
        """artery_network.py:565-665."""
        return self._dev.junction_eval(self._junction_of(p), [p.dex, d1.dex, d2.dex], x)[1]

    def newton(self, p, d1, d2, x=None, k_max=30, tol=1.e-10):
        """artery_network.py:668-710.  ``x`` is updated in place and returned."""
        if x is None:
            x = np.ones(18)
        sol, solves = self._dev.junction_newton(self._junction_of(p), [p.dex, d1.dex, d2.dex], x, k_max, tol)
        self.last_junction_solves = solves
        np.asarray(x)[...] = sol
        return x

    def adjust_bifurcation_step(self, p, d1, d2, margin=0.05):
        """artery_network.py:713-737."""
        dex = (1 + margin) * self.dt / self._dev.junction_cfl_min(self._junction_of(p))     # an:737, same order
        allx = self._dev.get_dex()
        for a in (p, d1, d2):
            allx[0, a._k] = dex
        self._dev.set_dex(allx)

    def set_inner_bc(self, ip, i1, i2):
        """artery_network.py:740-764."""
        p, d1, d2 = self.arteries[ip], self.arteries[i1], self.arteries[i2]
        self.adjust_bifurcation_step(p, d1, d2)
        k = self.range_parent_arteries.index(ip)
        x = self.x
        x[k] = self.newton(p, d1, d2, np.array(x[k]))
        p.U_out = [x[k, 9], x[k, 0]]
        d1.U_in = [x[k, 12], x[k, 3]]
        d2.U_in = [x[k, 15], x[k, 6]]

    def set_bcs(self, q_in):
        """artery_network.py:767-786: inlet value, then every junction and
        every leaf in ONE device launch (they only read Un)."""
        self.arteries[0].q_in = q_in
        self._dev.phase_boundary()

    # ---- output ---------------------------------------------------------------------
    def dump_metadata(self):
        """artery_network.py:789-854: mesh files + data.cfg.  Meshes are written
        as .npy vertex arrays (no HDF5 library in this environment)."""
        out = self.sol['output_location']
        os.makedirs(out, exist_ok=True)
        mesh_locations = []
        for a in self.arteries:
            if a is not None:
                loc = out + '/mesh_%i.npy' % a.i
                np.save(loc, a.dx * np.arange(a.Nx + 1))
                mesh_locations.append(loc)
        names, locations = ['flow'], [out + '/flow']
        if self.sol['store_area']:
            names.append('area'); locations.append(out + '/area')
        if self.sol['store_pressure']:
            names.append('pressure'); locations.append(out + '/pressure')
        N_cycles, N_cycles_store = self.geo['N_cycles'], self.sol['N_cycles_store']
        cfg = configparser.RawConfigParser()
        cfg.add_section('data')
        for key, value in (
                ('order', self.nondim['order']), ('Nx', self.geo['Nx']),
                ('Nt', self.sol['Nt_store'] * N_cycles_store), ('T0', self.T * (N_cycles - N_cycles_store)),
                ('T', self.T * N_cycles), ('L', ' '.join(repr(float(v)) for v in self.nondim['L'])),
                ('rc', self.nondim['rc']),
                ('qc', self.nondim['qc']), ('rho', self.nondim['rho']),
                ('mesh_locations', ','.join(mesh_locations)), ('names', ','.join(names)),
                ('locations', ','.join(locations))):
            cfg.set('data', key, str(value))
        with open(out + '/data.cfg', 'w') as fh:
            cfg.write(fh)

    def store_mask(self):
        """The dump test ``n % (Nt/Nt_store) == 0`` of artery_network.py:926,
        float modulo included, for every n of a cycle."""
        Nt, Nt_store = self.geo['Nt'], self.sol['Nt_store']
        # numpy's float remainder is Python's for non-negative operands (both are C fmod)
        return (np.arange(Nt) % (Nt / Nt_store) == 0).astype(np.uint8)

    def solve(self, write_output=True, n_steps=None):
        """artery_network.py:857-946.  The whole loop runs on the device; the
        host only enqueues it, then fetches the stored snapshots and writes
        them in the reference's directory layout.  Returns the snapshots as a
        dict (``t``, ``flow``, ``area``, ``pressure``: [snapshot, artery, vertex]).
        ``n_steps`` (not in the reference) stops the loop early."""
        self.define_x()
        self._dev.clear_flags()
        if write_output:
            self.dump_metadata()
        Nt, N_cycles = self.geo['Nt'], self.geo['N_cycles']
        N_cycles_store = self.sol['N_cycles_store']
        mask = self.store_mask()
        fields = _ffi.AF_STORE_FLOW
        if self.sol['store_area']:
            fields |= _ffi.AF_STORE_AREA
        if self.sol['store_pressure']:
            fields |= _ffi.AF_STORE_PRESSURE
        first_cycle = N_cycles - N_cycles_store
        total = N_cycles * Nt if n_steps is None else min(int(n_steps), N_cycles * Nt)
        # snapshots that fall into the executed range (the store is sized for exactly those)
        g = np.arange(total)
        n_snap = int(np.count_nonzero((g // Nt >= first_cycle) & (mask[g % Nt] != 0)))
        self._dev.set_store(mask, first_cycle, max(1, n_snap), fields)
        self._dev.stream_store_to_host()         # snapshots leave the device while the loop runs
        self._dev.step(0, total)
        snaps = self._dev.fetch_snapshots()      # streams the snapshots out while the loop is still running
        # a failed artery solve ends the loop at its step: the launches queued behind it return at once and
        # af_net_step stops enqueuing (DOLFIN raises inside that step, artery.py:244)
        bad = self._dev.failed_step()
        flags = int(self._dev.get_flags()[0])
        self.flags = flags
        if bad >= 0 or flags & _ffi.AF_FLAG_ARTERY_NOCONV:
            self.failed_step = bad
            raise RuntimeError("Newton solver did not converge in an artery solve at time step %d "
                               "(cycle %d, n = %d; flags=0x%x%s; DOLFIN raises here, artery.py:244)"
                               % (bad, bad // Nt, bad % Nt, flags,
                                  ", non-finite residual" if flags & _ffi.AF_FLAG_NONFINITE else ""))
        # t is accumulated by repeated "t += dt" (artery_network.py:946)
        t_all = np.concatenate([[0.0], np.cumsum(np.full(N_cycles * Nt, self.dt))])
        res = {'t': t_all[snaps['step']], 'step': snaps['step']}
        for name in ('flow', 'area', 'pressure'):
            if name in snaps:
                res[name] = snaps[name][:, 0]
        self.flags = flags
        if write_output:
            self._write_output(res)
        return res

    def _write_output(self, res):
        out = self.sol['output_location']
        for name in ('flow', 'area', 'pressure'):
            if name not in res:
                continue
            for a in self.arteries:
                if a is None:
                    continue
                f = _SeriesFile('%s/%s/%s_%i.xdmf' % (out, name, name, a.i))
                f.write_series(res[name][:, a._k], name, res['t'])
