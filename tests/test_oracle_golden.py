"""The oracle against the golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py): host set-up, every boundary-condition function
of the hot path, and the full time loop (reference loop on the live form, see make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import arteryfe_oracle as ao

CASES = ('c1', 'c2', 'o3')
BC_STEPS = (0, 3, 25)


def _net(case):
    return ao.network_from_cfg(os.path.join(GOLDEN, 'golden_%s.cfg' % case))


def _arts(net):
    return [a for a in net.arteries if a is not None]


@pytest.mark.parametrize('case', CASES)
def test_setup_matches_reference(case):
    g = np.load(os.path.join(GOLDEN, 'setup_%s.npz' % case))
    net = _net(case)
    for k in ('Ru', 'Rd', 'L', 'k1', 'k2', 'k3', 'Re', 'nu', 'p0', 'R1', 'R2', 'CT'):
        np.testing.assert_array_equal(np.asarray(net.nondim[k], dtype=float), g['nondim_' + k])
    assert net.T == g['T'] and net.dt == g['dt']
    np.testing.assert_array_equal(net.q_ins, g['q_ins'])
    np.testing.assert_array_equal(net.range_parent_arteries, g['range_parent'])
    np.testing.assert_array_equal(net.range_daughter_arteries, g['range_daughter'])
    np.testing.assert_array_equal(net.range_leaf_arteries, g['range_leaf'])
    arts = _arts(net)
    np.testing.assert_array_equal([a.dx for a in arts], g['dx'])
    np.testing.assert_array_equal([a.db for a in arts], g['db'])
    np.testing.assert_allclose([a.q0 for a in arts], g['q0'], rtol=1e-15)
    for key in ('R1', 'R2', 'CT'):
        np.testing.assert_array_equal(
            [net.arteries[i].param[key] for i in net.range_leaf_arteries], g['leaf_' + key])
    for name in ('r0', 'A0', 'f', 'dfdr', 'drdx'):
        got = np.array([[getattr(a, name)(s * a.L) for s in g['expr_frac']] for a in arts])
        np.testing.assert_allclose(got, g['expr_' + name], rtol=2e-15, atol=0)
    np.testing.assert_allclose([a._Un for a in arts], g['Un0'], rtol=2e-15)
    np.testing.assert_array_equal([a.pn_values for a in arts], g['pn0'])
    net.define_x()
    np.testing.assert_allclose(net.x, g['x0'], rtol=2e-15)


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('step', BC_STEPS)
def test_boundary_functions_match_reference(case, step):
    g = np.load(os.path.join(GOLDEN, 'bc_%s.npz' % case))
    tag = 's%d_' % step
    net = _net(case)
    arts = _arts(net)
    for a, Un in zip(arts, g[tag + 'Un']):
        a._Un[:] = Un
    # flux / source / compute_U_half / CFL_term samples
    for a, smp in zip((net.arteries[0], net.arteries[net.range_leaf_arteries[-1]]), g[tag + 'samples']):
        x0, x1 = smp[0], smp[1]
        U0, U1 = a.Un(x0), a.Un(x1)
        got = np.concatenate([[x0, x1], U0, U1, net.flux(a, U0, x0), net.source(a, U0, x0),
                              net.compute_U_half(a, x0, x1, U0, U1), [a.CFL_term(x1, U1[0], U1[1])]])
        np.testing.assert_allclose(got, smp, rtol=1e-13, atol=1e-13)
    for k, ip in enumerate(net.range_parent_arteries):
        i1, i2 = net.daughter_arteries(ip)
        p, d1, d2 = net.arteries[ip], net.arteries[i1], net.arteries[i2]
        net.adjust_bifurcation_step(p, d1, d2)
        np.testing.assert_allclose(p.dex, g[tag + 'junction_dex'][k], rtol=1e-14)
        x = g[tag + 'x_in'][k].copy()
        np.testing.assert_allclose(net.problem_function(p, d1, d2, x), g[tag + 'problem_function'][k],
                                   rtol=1e-12, atol=1e-11)
        np.testing.assert_allclose(net.jacobian(p, d1, d2, x), g[tag + 'jacobian'][k], rtol=1e-13, atol=0)
        np.testing.assert_allclose(net.newton(p, d1, d2, x.copy()), g[tag + 'newton'][k], rtol=1e-12)
    wk = [net.windkessel(net.arteries[i]) for i in net.range_leaf_arteries]
    np.testing.assert_allclose(wk, g[tag + 'windkessel'], rtol=1e-13)
    np.testing.assert_allclose([net.arteries[i].dex for i in net.range_leaf_arteries],
                               g[tag + 'leaf_dex'], rtol=1e-14)


@pytest.mark.parametrize('case', ('c1',))
def test_time_loop_matches_reference_loop(case):
    """Oracle solve() against the reference's own ArteryNetwork.solve() loop
    (BCs, dump rule, time accumulation by reference code; per-artery FEM solve
    by the oracle restatement)."""
    g = np.load(os.path.join(GOLDEN, 'run_%s.npz' % case))
    net = _net(case)
    res = net.solve(record_counters=True)
    np.testing.assert_array_equal(res['t'], g['t'])
    for field in ('flow', 'area', 'pressure'):
        np.testing.assert_allclose(res[field], g[field], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(net.x, g['x_final'], rtol=1e-10)
    np.testing.assert_array_equal(res['artery_its'], g['artery_its'])
