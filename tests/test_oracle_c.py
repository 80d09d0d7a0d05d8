"""The C oracle (oracle/oracle.c: CPU baseline and checker for larger cases)
against the numpy oracle, which is pinned against the reference's own code."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import arteryfe_oracle as ao
from oracle.oracle_c import COracle


@pytest.mark.parametrize('case,steps', (('c1', 60), ('o3', 30)))
def test_c_oracle_matches_numpy_oracle(case, steps):
    cfg = os.path.join(GOLDEN, 'golden_%s.cfg' % case)
    net = ao.network_from_cfg(cfg)
    res = net.solve(n_steps=steps, record_counters=True)
    c = COracle(ao.network_from_cfg(cfg))
    failed, la, lj = c.step(0, steps, nthreads=2, log=True)
    assert failed == 0
    A, q = c.state()
    for k, a in enumerate(net.arteries):
        np.testing.assert_allclose(A[k], a._Un[:, 0], rtol=1e-11)
        np.testing.assert_allclose(q[k], a._Un[:, 1], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(c.x(), net.x, rtol=1e-10)
    np.testing.assert_array_equal(la, res['artery_its'])
    np.testing.assert_array_equal(lj, res['junction_solves'])


def test_c_oracle_matches_reference_run():
    """Full golden run of case c2 (500 steps, reference's own loop)."""
    g = np.load(os.path.join(GOLDEN, 'run_c2.npz'))
    cfg = os.path.join(GOLDEN, 'golden_c2.cfg')
    net = ao.network_from_cfg(cfg)
    c = COracle(net)
    steps = np.round(g['t'] / net.dt).astype(int)
    done = 0
    for k, s in enumerate(steps):
        c.step(done, int(s - done), nthreads=1)
        done = int(s)
        A, q = c.state()
        np.testing.assert_allclose(q, g['flow'][k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(A, g['area'][k], rtol=1e-10)
        np.testing.assert_allclose(c.pressure(), g['pressure'][k], rtol=1e-9)


def test_fine_grid_trees_are_sensitive():
    """Documents a property of the REFERENCE SCHEME, not of any implementation: on
    the fine-grid synthetic trees of BASELINE.json (order 5, Nx=500, Nt=4000,
    theta=0.55) a relative perturbation of 1e-13 in one vertex grows by more than
    ten orders of magnitude within 320 steps (grid-scale oscillations fed by the
    explicit, dex-wide boundary stencils next to a weakly damped theta=0.55
    interior).  Hence long-horizon element-wise parity at 1e-9 is unattainable
    for any two implementations on those configs; the parity tests use short
    horizons there and full cycles on the stable cases (c1, c2, o3)."""
    from bloodflow_b200.synthetic import tree_param

    def make():
        p = tree_param(5, Nx=500, Nt=4000, theta=0.55)
        return COracle(ao.OracleNetwork(p.param, p.geo, p.solution))
    a, b = make(), make()
    A, q = b.state()
    A[3, 100] *= 1 + 1e-13
    b.set_state(A, q)
    growth = []
    done = 0
    for s in (40, 320):
        a.step(done, s - done, nthreads=0)
        b.step(done, s - done, nthreads=0)
        done = s
        growth.append(np.abs(b.state()[0] / a.state()[0] - 1).max())
    assert growth[0] < 1e-10          # still tracking after 40 steps
    assert growth[1] > 1e-4           # decorrelated after 320


def test_theta_06_trees_are_not_sensitive():
    """The same tree with theta = 0.6 (the synthetic default, what bench.py runs): the perturbation
    stays at round-off level.  (Whole cycles / 4 cycles of C3 and C4: profiles/r02_stability_*.jsonl.)"""
    from bloodflow_b200.synthetic import tree_param

    def make():
        p = tree_param(5, Nx=500, Nt=4000, theta=0.6)
        return COracle(ao.OracleNetwork(p.param, p.geo, p.solution))
    a, b = make(), make()
    A, q = b.state()
    A[3, 100] *= 1 + 1e-13
    b.set_state(A, q)
    a.step(0, 400, nthreads=0)
    b.step(0, 400, nthreads=0)
    assert np.abs(b.state()[0] / a.state()[0] - 1).max() < 1e-11
