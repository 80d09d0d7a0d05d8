"""The per-artery FEM half against fixtures made from the REFERENCE'S OWN FORM TEXT.

``tests/golden/fem_<case>.npz`` hold, at steps 0/3/25 of the reference's time loop,
the residual vector and the block-tridiagonal Jacobian of every artery before and
after the Dirichlet rows, the first two Newton updates and the converged iterate.
They were produced by ``tests/golden/make_golden.py``: the unmodified
``/root/reference/arteryfe/artery.py:191-230`` builds ``variational_form`` out of
live UFL symbols, ``derivative(F, U)`` (``artery.py:238``) is a symbolic Gateaux
derivative, and a generic P1 / 3-point-Gauss assembler integrates it
(``tests/golden/ufl_live.py``).  Both oracles (numpy and C) are held to them here;
the CUDA kernel in ``test_gpu_parity.py::test_artery_solve_matches_reference_form``."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import arteryfe_oracle as ao
from oracle import oracle_c

CASES = ('c1', 'c2', 'o3')
STEPS = (0, 3, 25)


def _load(case):
    return np.load(os.path.join(GOLDEN, 'fem_%s.npz' % case))


def _set_bc(a, bc4):
    if a.root:
        a.q_in = bc4[1]
    else:
        a.U_in = np.array(bc4[0:2])
    if a.leaf:
        a.A_out = bc4[2]
    else:
        a.U_out = np.array(bc4[2:4])


def _close(x, ref, tol, scale=None):
    """Element-wise within tol of ``scale`` (default: the largest magnitude of the quantity)."""
    scale = np.abs(ref).max() if scale is None else scale
    np.testing.assert_allclose(x, ref, rtol=0, atol=tol * scale)


def _term_scale(g, tag, k):
    """Magnitude of the integrated terms that cancel inside a residual entry (|dt F2 phi'| ~ 10-30
    while |R| ~ 1e-2): |J| |U|.  Residuals are compared to 1e-13 of THIS, i.e. to ~1e-12 absolute."""
    return np.abs(g[tag + 'J0_raw'][k]).max() * np.abs(g[tag + 'U0'][k]).max()


@pytest.mark.parametrize('case', CASES)
def test_numpy_oracle_assembly_matches_reference_form(case):
    g = _load(case)
    net = ao.network_from_cfg(os.path.join(GOLDEN, 'golden_%s.cfg' % case))
    arts = [a for a in net.arteries if a is not None]
    for s in STEPS:
        for k, a in enumerate(arts):
            tag = 's%d_' % s
            a._Un[:] = g[tag + 'U0'][k]
            _set_bc(a, g[tag + 'bc4'][k])
            for it in ('0', '1'):
                U = g[tag + 'U' + it][k]
                if np.isnan(U).any():
                    continue
                ts, us = _term_scale(g, tag, k), np.abs(U).max()
                Rraw, Jraw = a.assemble(U, jacobian=True, apply_bc=False)
                _close(Rraw, g[tag + 'R%s_raw' % it][k], 1e-13, ts)
                R, J = a.assemble(U, jacobian=True)
                _close(R, g[tag + 'R' + it][k], 1e-13, ts)
                if not np.isnan(g[tag + 'J%s_raw' % it][k]).any():
                    _close(np.array(Jraw), g[tag + 'J%s_raw' % it][k], 1e-13)
                    _close(np.array(J), g[tag + 'J' + it][k], 1e-13)
                    dU = a._block_tridiag_solve(J, R)
                    _close(dU, g[tag + 'dU' + it][k], 1e-11, us)      # i.e. the next iterate to 1e-11
            # the whole solve: same number of linear solves, same iterate
            a._U[:] = g[tag + 'U0'][k]
            a.solve()
            assert a.newton_its == g[tag + 'its'][k]
            np.testing.assert_allclose(a._U, g[tag + 'U_final'][k], rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('case', CASES)
def test_c_oracle_assembly_matches_reference_form(case):
    g = _load(case)
    net = ao.network_from_cfg(os.path.join(GOLDEN, 'golden_%s.cfg' % case))
    co = oracle_c.COracle(net)
    for s in STEPS:
        tag = 's%d_' % s
        U0 = g[tag + 'U0']
        co.set_state(U0[:, :, 0], U0[:, :, 1])
        bc = np.nan_to_num(g[tag + 'bc4'])
        co.set_bc(bc)
        for k in range(co.na):
            for it in ('0', '1'):
                U = g[tag + 'U' + it][k]
                if np.isnan(U).any() or np.isnan(g[tag + 'J' + it][k]).any():
                    continue
                R, J = co.artery_assemble(k, U[:, 0], U[:, 1])
                _close(R, g[tag + 'R' + it][k], 1e-13, _term_scale(g, tag, k))
                _close(J, g[tag + 'J' + it][k], 1e-13)
            # one linear solve, then the whole solve
            assert co.artery_solve(k, 1) == min(1, g[tag + 'its'][k])
            A, q = co.pending()
            np.testing.assert_allclose(A[k], g[tag + 'U1'][k][:, 0], rtol=1e-11)
            np.testing.assert_allclose(q[k], g[tag + 'U1'][k][:, 1], rtol=1e-11, atol=1e-13)
        co.set_state(U0[:, :, 0], U0[:, :, 1])
        for k in range(co.na):
            assert co.artery_solve(k) == g[tag + 'its'][k]
        A, q = co.pending()
        np.testing.assert_allclose(A, g[tag + 'U_final'][:, :, 0], rtol=1e-11)
        np.testing.assert_allclose(q, g[tag + 'U_final'][:, :, 1], rtol=1e-11, atol=1e-13)
    co.close()
