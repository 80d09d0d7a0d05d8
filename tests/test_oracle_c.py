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
    """Full hybrid golden run of case c2 (500 steps, reference's own loop)."""
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
