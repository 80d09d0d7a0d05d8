"""Host-side helpers with the names and semantics of the reference's
arteryfe/utils.py: units, nondimensionalisation, inlet table, output reader.
Everything here is set-up or post-processing; none of it is on the GPU path."""
import configparser
import copy
import os

import numpy as np
from scipy.interpolate import interp1d

__all__ = ['unit_to_mmHg', 'mmHg_to_unit', 'nondimensionalise_parameters', 'nondimensionalise',
           'redimensionalise', 'read_inlet', 'read_output', 'XDMF_to_matrix', 'plot_matrix', 'is_near',
           'write_file', 'read_file', 'print_progress']

_MMHG = 76 / 101325       # utils.py:13-44


def unit_to_mmHg(p):
    """g cm^-1 s^-2 -> mmHg."""
    return _MMHG * p


def mmHg_to_unit(p):
    """mmHg -> g cm^-1 s^-2."""
    return 101325 / 76 * p


def nondimensionalise_parameters(params):
    """utils.py:47-104.  ``R_term`` and ``p_term`` may be absent (the shipped
    configs lack them); they are carried through when present."""
    p = params.param
    nd = copy.deepcopy(p)
    rc, rho, qc = p['rc'], p['rho'], p['qc']
    for key in ('Ru', 'Rd', 'L'):
        nd[key] = p[key] / rc
    if 'R_term' in p:
        nd['R_term'] = p['R_term'] / rc
    nd['k1'] = p['k1'] * rc ** 4 / rho / qc ** 2
    nd['k2'] = p['k2'] * rc
    nd['k3'] = p['k3'] * rc ** 4 / rho / qc ** 2
    nd['Re'] = p['qc'] / p['nu'] / rc
    nd['nu'] = p['nu'] * rc / qc
    nd['p0'] = p['p0'] * rc ** 4 / rho / qc ** 2
    if 'p_term' in p:
        nd['p_term'] = p['p_term'] * rc ** 4 / rho / qc ** 2
    nd['R1'] = p['R1'] * rc ** 4 / rho / qc
    nd['R2'] = p['R2'] * rc ** 4 / rho / qc
    nd['CT'] = p['CT'] * rho * qc ** 2 / rc ** 7
    return nd


def nondimensionalise(rc, qc, rho, x, nature):
    if nature == 'time':
        return x * qc / rc ** 3
    if nature == 'area':
        return x / rc ** 2
    if nature == 'flow':
        return x / qc
    if nature == 'pressure':
        return x * rc ** 4 / rho / qc ** 2
    return x


def redimensionalise(rc, qc, rho, x, nature):
    if nature == 'time':
        return x * rc ** 3 / qc
    if nature == 'area':
        return x * rc ** 2
    if nature == 'flow':
        return x * qc
    if nature == 'pressure':
        return x * rho * qc ** 2 / rc ** 4
    return x


def read_inlet(data_location, Nt):
    """utils.py:173-200: (T, q_ins) with q_ins the CSV linearly interpolated at
    ``linspace(0, T, Nt)`` -- spacing T/(Nt-1), not dt."""
    key = (os.path.abspath(data_location), os.path.getmtime(data_location), int(Nt))
    hit = _inlet_cache.get(key)
    if hit is None:                      # parameter studies build many networks on the same inlet file
        table = np.genfromtxt(data_location, delimiter=',')
        T = table[-1, 0]
        q_ins = interp1d(table[:, 0], table[:, 1])(np.linspace(0, T, Nt))
        if len(_inlet_cache) > 16:
            _inlet_cache.clear()
        hit = _inlet_cache[key] = (T, q_ins)
    return hit[0], hit[1].copy()


_inlet_cache = {}


def read_output(filename):
    """utils.py:203-234: the data.cfg written by ArteryNetwork.dump_metadata."""
    cfg = configparser.ConfigParser()
    cfg.read(filename)
    g = lambda k: cfg.get('data', k)      # noqa: E731
    L = [float(tok) for tok in g('L').replace(',', ' ').split()]
    return (cfg.getint('data', 'order'), cfg.getint('data', 'Nx'), cfg.getint('data', 'Nt'),
            cfg.getfloat('data', 'T0'), cfg.getfloat('data', 'T'), L, cfg.getfloat('data', 'rc'),
            cfg.getfloat('data', 'qc'), cfg.getfloat('data', 'rho'), g('mesh_locations').split(','),
            g('names').split(','), g('locations').split(','))


def XDMF_to_matrix(Nx, Nt, mesh_location, location, name):
    """utils.py:237-272: M[Nx+1, Nt] in ascending x.  Reads the XDMF files this
    package writes (heavy data in a raw little-endian float64 sidecar; DOLFIN's
    HDF5 checkpoint layout is not reproduced -- no HDF5 library exists here)."""
    side = os.path.splitext(location)[0] + '.bin'
    raw = np.fromfile(side, dtype='<f8')
    M = raw.reshape(-1, Nx + 1)[:Nt].T.copy()
    assert M.shape == (Nx + 1, Nt)
    return M


def plot_matrix(t, x, M, label, output):
    """utils.py:275-302 (needs matplotlib, which is optional here)."""
    import matplotlib
    matplotlib.use('Agg')
    import matplotlib.pyplot as plt
    T, X = np.meshgrid(t, x)
    fig = plt.figure(figsize=(8, 6))
    ax = fig.add_subplot(projection='3d')
    ax.plot_surface(T, X, M, rstride=1, cstride=1, cmap='viridis', linewidth=0, antialiased=False)
    ax.set_xlabel('t'); ax.set_ylabel('x'); ax.set_zlabel(label)
    ax.set_ylim(min(x), max(x)); ax.set_xlim(min(t), max(t))
    plt.savefig(output)


def is_near(a, b, tol=1.e-11, reltol=1.e-10):
    """utils.py:305-330."""
    if np.abs(b) > 1.e-10:
        return bool(np.abs(a - b) < tol or np.abs(a / b - 1) < reltol)
    return bool(np.abs(a - b) < tol)


class _SeriesFile(object):
    """Stand-in for dolfin.XDMFFile as used by write_file/read_file: an XDMF
    XML index plus a raw float64 sidecar, one row of Nx+1 values per write."""

    def __init__(self, path):
        self.path = path
        self.side = os.path.splitext(path)[0] + '.bin'
        self.times = []
        os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
        open(self.side, 'wb').close()

    def write_checkpoint(self, values, label, t):
        with open(self.side, 'ab') as fh:
            np.asarray(values, dtype='<f8').tofile(fh)
        self.times.append((label, float(t), len(values)))
        self._write_index()

    def write_series(self, rows, label, times):
        """All snapshots of one artery at once: rows [n_snap, Nx+1]."""
        rows = np.ascontiguousarray(rows, dtype='<f8')
        with open(self.side, 'ab') as fh:
            rows.tofile(fh)
        self.times += [(label, float(t), rows.shape[1]) for t in times]
        self._write_index()

    def read_checkpoint(self, label, i):
        n = self.times[i][2] if self.times else None
        raw = np.fromfile(self.side, dtype='<f8')
        return raw.reshape(-1, n)[i] if n else raw

    def _write_index(self):
        rows = []
        for k, (label, t, n) in enumerate(self.times):
            rows.append(
                '      <Grid Name="%s_%d"><Time Value="%.17g"/>'
                '<Attribute Name="%s" Center="Node"><DataItem Format="Binary" NumberType="Float" Precision="8" '
                'Endian="Little" Seek="%d" Dimensions="%d">%s</DataItem></Attribute></Grid>'
                % (label, k, t, label, 8 * n * k, n, os.path.basename(self.side)))
        with open(self.path, 'w') as fh:
            fh.write('<?xml version="1.0"?>\n<Xdmf Version="3.0">\n  <Domain>\n'
                     '    <Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">\n'
                     + '\n'.join(rows) + '\n    </Grid>\n  </Domain>\n</Xdmf>\n')


def write_file(f, u, label, t):
    """utils.py:333-350."""
    f.write_checkpoint(u, label, t)


def read_file(f, u, label, i):
    """utils.py:353-370: fills the array ``u`` in place."""
    u[...] = f.read_checkpoint(label, i)


def print_progress(n_cycle, n, dt):
    """utils.py:373-387."""
    print('Current cycle: %i, Cycle iteration: %i, Time-step %i' % (n_cycle, n, dt), end='\r')
