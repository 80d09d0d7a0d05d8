"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The unmodified reference modules are imported under ``dolfin_standin`` (see its
docstring).  All boundary-condition arithmetic, the time loop, the parameter
handling and the initial conditions that end up in the fixtures are computed by
reference code (``arteryfe/artery_network.py``, ``artery.py``, ``utils.py``,
``param_parser.py``).  FEniCS' per-artery variational solve ([3P], un-vendored)
is performed on the reference's OWN ``variational_form`` / ``derivative(F, U)``
objects (artery.py:191-238) by the generic assembler + NewtonSolver loop of
``ufl_live.py``; the repo's oracle is not involved in producing any fixture.

Fixtures (numpy .npz, a few hundred kB in total):

* setup_<case>.npz     nondimensional parameters, T, dt, q_ins, per-artery
                       constants, Expression samples, initial state, x0
* bc_<case>.npz        at steps k in BC_STEPS of the reference loop: Un of every
                       artery, dex, and for every junction problem_function(x),
                       jacobian(x), newton(x) / for every leaf windkessel();
                       plus flux/source/compute_U_half/CFL_term samples
* fem_<case>.npz       at the same steps, per artery: iterate U0 (= Un), Dirichlet
                       values, residual and block-tridiagonal Jacobian before
                       (``*_raw``) and after the Dirichlet rows at the first two
                       Newton iterates, the updates dU0/dU1, the converged
                       iterate and the number of linear solves
* run_<case>.npz       stored flow/area/pressure of the reference's own
                       ArteryNetwork.solve() loop, with times and Newton counts
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = '/root/reference'
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import dolfin_standin  # noqa: E402

dolfin_standin.install()
sys.path.insert(0, REFERENCE)

import arteryfe as ref  # noqa: E402  (the reference package)

INLET = os.path.join(ROOT, 'data', 'example_inlet.csv')

PHYS = """k1 = 2.0e7
k2 = -22.53
k3 = 8.65e5
rho = 1.06
nu = 0.046
p0 = 119990.131579
"""

# The shipped configs do not load against HEAD: the reference needs
# R_term, p_term and per-leaf R1/R2/CT arrays.  These cases add exactly those
# keys and shrink Nx/Nt so the fixtures stay small.
CASES = {
    # demo_arterybranch.cfg geometry (tapered daughters); Windkessel values
    # chosen so that HEAD's 0.3/0.01 factors give a contracting Picard map
    'c1': dict(params="""order = 2
rc = 1.0
qc = 10.0
Ru = 0.37,0.177,0.177
Rd = 0.37,0.17,0.17
L = 20.8,17.7,17.6
R_term = 0.17
p_term = 6000
R1 = 84333.3,84333.3
R2 = 1390000.0,1390000.0
CT = 1.3384e-6,1.3384e-6
""", Nx=40, Nt=400, N_cycles=1, Nt_store=40, theta=0.55),
    # demo_arterybranch_geo.cfg + R_term (alpha geometry, untapered leaves).  The
    # shipped R1/R2 make the explicit Windkessel update unstable at this coarse
    # Nt (dt/(R2_eff*CT) > 2), so c1-like effective values are used instead.
    'c2': dict(params="""order = 2
rc = 1.0
qc = 10.0
Ru = 0.37,1,1
Rd = 0.37,1,1
alpha= 0.8
L = 20
R_term = 0.296
p_term = 6000
R1 = 84333.3,80000.0
R2 = 1390000.0,1300000.0
CT = 1.3384e-6,1.5e-6
""", Nx=50, Nt=500, N_cycles=1, Nt_store=50, theta=0.55),
    # order-3 synthetic tree, non-unit rc/qc to exercise the scalings,
    # per-leaf Windkessel values all different, around the SURVEY 8(d) recipe
    # (R1_eff = 2 Z0; larger R1/Z0 ratios blow this scheme up within a cycle)
    'o3': dict(params="""order = 3
rc = 1.1
qc = 9.0
Ru = 0.37,1,1,1,1,1,1
Rd = 0.37,1,1,1,1,1,1
alpha= 0.8
L = 20
R_term = 0.2368
p_term = 6000
R1 = 31192.6,31500.0,30900.0,32000.0
R2 = 7.79e6,7.6e6,7.9e6,7.7e6
CT = 2.387e-7,2.4e-7,2.3e-7,2.5e-7
""", Nx=32, Nt=640, N_cycles=1, Nt_store=32, theta=0.55),
}
BC_STEPS = (0, 3, 25)


def write_cfg(case, path, outdir):
    c = CASES[case]
    with open(path, 'w') as fh:
        fh.write('[Parameters]\n' + c['params'] + PHYS)
        fh.write('\n[Geometry]\nNx = %d\nNt = %d\nN_cycles = %d\n' % (c['Nx'], c['Nt'], c['N_cycles']))
        fh.write('\n[Solution]\ninlet_flow_location = %s\noutput_location = %s\n' % (INLET, outdir))
        fh.write('theta = %r\nNt_store = %d\nN_cycles_store = 1\nstore_area = 1\nstore_pressure = 1\n'
                 % (c['theta'], c['Nt_store']))


def reference_network(cfg):
    argv = sys.argv
    sys.argv = ['make_golden', '--cfg', cfg]
    try:
        param = ref.ParamParser()
    finally:
        sys.argv = argv
    return ref.ArteryNetwork(param)


class SolveLog(object):
    """Observer of the stand-in's NonlinearVariationalSolver.solve(): logs the
    number of linear solves of every per-artery Newton solve."""

    def __init__(self):
        self.its = []

    def __call__(self, problem, its):
        self.its.append(its)


def arteries_of(net):
    return [a for a in net.arteries if a is not None]


def golden_setup(net, case):
    d = {}
    nd = net.nondim
    for k in ('Ru', 'Rd', 'L', 'k1', 'k2', 'k3', 'Re', 'nu', 'p0', 'R1', 'R2', 'CT', 'R_term', 'p_term'):
        d['nondim_' + k] = np.asarray(nd[k], dtype=float)
    d['T'], d['dt'], d['q_ins'] = net.T, net.dt, net.q_ins
    d['range_parent'] = np.array(net.range_parent_arteries)
    d['range_daughter'] = np.array(net.range_daughter_arteries)
    d['range_leaf'] = np.array(net.range_leaf_arteries)
    arts = arteries_of(net)
    d['dx'] = np.array([a.dx for a in arts])
    d['db'] = np.array([a.db for a in arts])
    d['q0'] = np.array([a.q0 for a in arts])
    d['leaf_R1'] = np.array([net.arteries[i].param['R1'] for i in net.range_leaf_arteries])
    d['leaf_R2'] = np.array([net.arteries[i].param['R2'] for i in net.range_leaf_arteries])
    d['leaf_CT'] = np.array([net.arteries[i].param['CT'] for i in net.range_leaf_arteries])
    # Expression samples, inside and outside the vessel (the BC code evaluates
    # them at L+dex and -dex)
    fr = np.array([-0.07, 0.0, 0.21, 0.5, 0.83, 1.0, 1.04])
    d['expr_frac'] = fr
    for name in ('r0', 'A0', 'f', 'dfdr', 'drdx'):
        d['expr_' + name] = np.array([[getattr(a, name)(s * a.param['L']) for s in fr] for a in arts])
    d['Un0'] = np.array([a.Un.values for a in arts])
    d['pn0'] = np.array([a.pn.values[:, 0] for a in arts])
    net.define_x()
    d['x0'] = net.x.copy()
    return d


def golden_bc(net):
    """Reference BC functions evaluated on the states of a hybrid run."""
    d, fem = {}, {}
    Nt = net.geo['Nt']
    arts = arteries_of(net)
    net.define_x()                      # what solve() does first (artery_network.py:874)
    for n in range(max(BC_STEPS) + 1):
        if n in BC_STEPS:
            tag = 's%d_' % n
            d[tag + 'Un'] = np.array([a.Un.values for a in arts])
            # low-level samples on the parent artery (tapered in c1) and a leaf
            smp = []
            for a in (net.arteries[0], net.arteries[net.range_leaf_arteries[-1]]):
                L = a.param['L']
                x0, x1 = 0.37 * L, 0.37 * L + 3.3 * a.dx
                U0, U1 = a.Un(x0), a.Un(x1)
                smp.append(np.concatenate([
                    [x0, x1], U0, U1, net.flux(a, U0, x0), net.source(a, U0, x0),
                    net.compute_U_half(a, x0, x1, U0, U1), [a.CFL_term(x1, U1[0], U1[1])]]))
            d[tag + 'samples'] = np.array(smp)
            fx, Jx, xs, xin, dexs = [], [], [], [], []
            for k, ip in enumerate(net.range_parent_arteries):
                i1, i2 = net.daughter_arteries(ip)
                p, d1, d2 = net.arteries[ip], net.arteries[i1], net.arteries[i2]
                net.adjust_bifurcation_step(p, d1, d2)
                dexs.append(p.dex)
                x = net.x[k].copy()
                xin.append(x.copy())
                fx.append(net.problem_function(p, d1, d2, x))
                Jx.append(net.jacobian(p, d1, d2, x))
                xs.append(net.newton(p, d1, d2, x.copy()))
            if n == BC_STEPS[1]:
                # the 'Singular' branch of newton() (artery_network.py:701-708): with A^{n+1} of the three
                # vessels at 1e300 the pressure-continuity rows 10, 11 of the Jacobian underflow to zero, LAPACK
                # reports an exactly singular matrix and the reference perturbs (J + 1e-6 I, func[0] += 1e-6)
                ip = net.range_parent_arteries[0]
                i1, i2 = net.daughter_arteries(ip)
                p, d1, d2 = net.arteries[ip], net.arteries[i1], net.arteries[i2]
                x = xin[0].copy()
                x[[9, 12, 15]] = 1.0e300
                d['singular_x_in'] = x.copy()
                d['singular_dex'] = np.array([p.dex, d1.dex, d2.dex])
                said = io.StringIO()
                with contextlib.redirect_stdout(said), np.errstate(all='ignore'):
                    d['singular_newton_k1'] = net.newton(p, d1, d2, x.copy(), k_max=1)
                assert 'Singular' in said.getvalue()
            d[tag + 'junction_dex'] = np.array(dexs)
            d[tag + 'x_in'] = np.array(xin)
            d[tag + 'problem_function'] = np.array(fx)
            d[tag + 'jacobian'] = np.array(Jx)
            d[tag + 'newton'] = np.array(xs)
            wk, wdex = [], []
            for i in net.range_leaf_arteries:
                wk.append(net.windkessel(net.arteries[i]))
                wdex.append(net.arteries[i].dex)
            d[tag + 'windkessel'] = np.array(wk)
            d[tag + 'leaf_dex'] = np.array(wdex)
        # advance one step exactly like the loop body artery_network.py:916-944
        net.set_bcs(net.q_ins[(n + 1) % Nt])
        for k, a in enumerate(arts):
            if n in BC_STEPS:
                rec = dolfin_standin.record_next = {}
            a.solve()
            if n in BC_STEPS:
                bc4 = np.full(4, np.nan)            # A_in, q_in, A_out, q_out (NaN: no Dirichlet row)
                for node, comp, g in rec.pop('bc_rows'):
                    bc4[(0 if node == 0 else 2) + int(comp)] = g
                rec['bc4'] = bc4
                for missing, like in (('J1_raw', 'J0_raw'), ('J1', 'J0'), ('dU1', 'dU0'),
                                      ('U1', 'U0'), ('R1_raw', 'R0'), ('R1', 'R0')):
                    if missing not in rec:          # the solve stopped before that iterate existed
                        rec[missing] = np.full_like(rec[like], np.nan)
                for key, val in rec.items():
                    fem.setdefault('s%d_%s' % (n, key), []).append(val)
            a.update_solution()
    fem = {key: np.array(val) for key, val in fem.items()}
    return d, fem


def golden_run(net, log):
    dolfin_standin.XDMFFile.written.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        net.solve()
    out = net.sol['output_location']
    d = {}
    for field in ('flow', 'area', 'pressure'):
        rows = []
        for a in arteries_of(net):
            rec = dolfin_standin.XDMFFile.written['%s/%s/%s_%i.xdmf' % (out, field, field, a.i)]
            rows.append(np.array([v for _, v in rec]))
            d['t'] = np.array([t for t, _ in rec])
        d[field] = np.swapaxes(np.array(rows), 0, 1)       # [snapshot, artery, node]
    d['x_final'] = net.x.copy()
    d['artery_its'] = np.array(log.its).reshape(-1, len(arteries_of(net)))
    return d


def main():
    tmp = tempfile.mkdtemp(prefix='golden_')
    for case in CASES:
        cfg = os.path.join(HERE, 'golden_%s.cfg' % case)
        write_cfg(case, cfg, tmp)
        # keep the committed cfg free of container paths
        text = open(cfg).read()
        np.savez_compressed(os.path.join(HERE, 'setup_%s.npz' % case),
                            **golden_setup(reference_network(cfg), case))
        net = reference_network(cfg)
        dolfin_standin.solve_observer = None
        bc, fem = golden_bc(net)
        np.savez_compressed(os.path.join(HERE, 'bc_%s.npz' % case), **bc)
        np.savez_compressed(os.path.join(HERE, 'fem_%s.npz' % case), **fem)
        net = reference_network(cfg)
        log = dolfin_standin.solve_observer = SolveLog()
        np.savez_compressed(os.path.join(HERE, 'run_%s.npz' % case), **golden_run(net, log))
        open(cfg, 'w').write(text.replace(INLET, 'data/example_inlet.csv').replace(tmp, 'output/golden'))
        print('wrote fixtures for', case)


if __name__ == '__main__':
    main()
