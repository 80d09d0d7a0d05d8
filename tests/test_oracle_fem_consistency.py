"""Internal consistency of the restated FEniCS half ([3P], parity unpinned): the
Jacobian is the exact derivative of the quadrature-discretised residual (what
UFL's derivative() produces, artery.py:238), rest is a steady state of an
untapered vessel, the P2 coefficient interpolation is exact for quadratics, and
the Newton control follows dolfin::NewtonSolver (count = linear solves)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import arteryfe_oracle as ao


def _artery(case, idx):
    net = ao.network_from_cfg(os.path.join(GOLDEN, 'golden_%s.cfg' % case))
    return net, net.arteries[idx]


def _dense(blocks):
    Lo, Dg, Up = blocks
    n = Dg.shape[0]
    J = np.zeros((2 * n, 2 * n))
    for i in range(n):
        J[2 * i:2 * i + 2, 2 * i:2 * i + 2] = Dg[i]
        if i > 0:
            J[2 * i:2 * i + 2, 2 * i - 2:2 * i] = Lo[i]
        if i < n - 1:
            J[2 * i:2 * i + 2, 2 * i + 2:2 * i + 4] = Up[i]
    return J


@pytest.mark.parametrize('case,idx', (('c1', 0), ('c1', 1), ('o3', 2), ('o3', 6)))
def test_jacobian_is_derivative_of_residual(case, idx):
    net, a = _artery(case, idx)
    rng = np.random.default_rng(3)
    U = a._U * (1 + 0.03 * rng.standard_normal(a._U.shape))      # away from Un, generic state
    R, blocks = a.assemble(U, jacobian=True)
    J = _dense(blocks)
    n = U.size
    Jfd = np.zeros((n, n))
    for k in range(n):
        h = 1e-6 * max(abs(U.flat[k]), 1e-3)
        Up_, Um_ = U.copy(), U.copy()
        Up_.flat[k] += h
        Um_.flat[k] -= h
        Jfd[:, k] = (a.assemble(Up_) - a.assemble(Um_)).reshape(-1) / (2 * h)
    scale = np.abs(J).max()
    np.testing.assert_allclose(J, Jfd, rtol=0, atol=2e-7 * scale)
    # block-tridiagonal: nothing outside the three block diagonals
    mask = np.abs(np.subtract.outer(np.arange(n) // 2, np.arange(n) // 2)) > 1
    assert np.abs(Jfd[mask]).max() < 1e-9 * scale


def test_rest_state_is_stationary_on_untapered_vessel():
    """A = A0, q = 0 on a vessel with Rd == Ru: F2 is constant, S2 = 0, so the
    weak form residual vanishes in the interior and Newton needs no iteration."""
    net, a = _artery('c2', 0)            # root of the _geo tree: untapered
    a._Un[:, 1] = 0.0
    a._U[:] = a._Un
    a.q_in = 0.0
    a.U_out = np.array([a.A0(a.L), 0.0])
    R = a.assemble(a._U)
    # the end rows carry the (wrong-signed at x=0, ar:197-198) facet terms where not Dirichlet
    assert np.abs(R[1:-1]).max() < 1e-12 * a.f(0) * a.A0(0)
    a.solve()
    assert a.newton_its <= 1
    np.testing.assert_allclose(a._U[:, 0], a._Un[:, 0], rtol=1e-12)
    assert np.abs(a._U[:, 1]).max() < 1e-9


def test_p2_interpolant_at_gauss_points_is_exact_for_quadratics():
    net, a = _artery('c1', 1)
    xm = 0.5 * (a.xn[:-1] + a.xn[1:])
    poly = lambda x: 0.3 - 1.7 * x + 0.05 * x * x          # noqa: E731
    xi = ao.GAUSS_XI
    n0, n1, nm = (1 - xi) * (1 - 2 * xi), xi * (2 * xi - 1), 4 * xi * (1 - xi)
    got = (poly(a.xn[:-1])[:, None] * n0 + poly(a.xn[1:])[:, None] * n1 + poly(xm)[:, None] * nm)
    xq = a.xn[:-1, None] + a.dx * xi[None, :]
    np.testing.assert_allclose(got, poly(xq), rtol=1e-13)
    assert abs(ao.GAUSS_W.sum() - 1) < 1e-15
    # 3-point Gauss integrates degree 5 exactly on [0, 1]
    for k in range(6):
        assert abs((ao.GAUSS_W * xi ** k).sum() - 1.0 / (k + 1)) < 1e-15


def test_newton_control_follows_dolfin():
    """Count = number of linear solves; converged iff r/r0 < 1e-9 or r < 1e-10;
    the test is applied before the first solve; max_it raises RuntimeError."""
    net, a = _artery('c1', 0)
    net.define_x()
    net.set_bcs(net.q_ins[1])
    a.solve()
    its, r = a.newton_its, a.last_residual
    assert 1 <= its <= 3 and r < 1e-10
    a.solve()                              # already converged: zero linear solves
    assert a.newton_its == 0
    a.update_solution()
    net.set_bcs(net.q_ins[2])
    with pytest.raises(RuntimeError):
        a.solve(max_it=0)
    a._U[:] = a._Un
    a.solve(atol=1e-3, rtol=1e-3)          # loose absolute tolerance stops after the first step
    assert a.newton_its <= 1
