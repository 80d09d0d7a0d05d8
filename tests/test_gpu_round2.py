"""GPU tests added in round 2: the C5-shaped ensemble against the C oracle, the
reference's 'Singular' junction branch, failure at the failing step, plain stream
order against programmatic dependent launch, flag handling."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _facade(case, **kw):
    import bloodflow_b200.arteryfe as af
    return af.ArteryNetwork(af.ParamParser(os.path.join(GOLDEN, 'golden_%s.cfg' % case)), **kw)


def test_c5_shaped_ensemble_members_match_c_oracle():
    """BASELINE config 5's shape (order 4, Nx=100, Nt=2000, perturbed members) with enough members
    (640 x 15 = 9 600 boundary tasks >= 148 x 64) that the thread-per-task boundary kernel
    k_boundary<false>, the one-warp artery kernel (T = 32) and two member shares on sub-streams run;
    members first / 1 / middle / last against the C oracle of the same perturbed parameters for 200
    steps: state to 1e-9, artery Newton counts equal at every step, junction counts up to near-ties."""
    from bloodflow_b200 import ensemble as en
    from bloodflow_b200.synthetic import tree_param
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle
    base = tree_param(4, Nx=100, Nt=2000, N_cycles=2, Nt_store=100)
    first, M, steps = 1000, 640, 200
    ens = en.EnsembleNetwork(base, M, first_index=first)
    dev = ens.dev
    assert dev.nb * (dev.nj + dev.nl) >= 148 * 64
    dev.set_counter_log(steps)
    dev.step(0, steps)
    A, q = dev.get_state()
    art, jun, _ = dev.fetch_counter_log()
    for m in (0, 1, M // 2, M - 1):
        p = en.perturb_param(base, first + m)
        c = COracle(ao.OracleNetwork(p.param, p.geo, p.solution))
        failed, la, lj = c.step(0, steps, nthreads=0, log=True)
        assert not failed
        Ao, qo = c.state()
        np.testing.assert_allclose(A[m], Ao, rtol=1e-9, err_msg='member %d' % m)
        np.testing.assert_allclose(q[m], qo, rtol=1e-9, atol=1e-12, err_msg='member %d' % m)
        np.testing.assert_array_equal(art[:, m], la)
        # junction counts: equal up to near-ties of ||f|| < 1e-10 (see test_full_size_trees_full_cycle)
        assert np.count_nonzero(jun[:, m] != lj) <= 2 and np.abs(jun[:, m] - lj).max() <= 1
        c.close()
    assert ens.flagged_members() == 0


@pytest.mark.parametrize('case', ('c1', 'c2', 'o3'))
def test_singular_junction_branch_matches_reference(case):
    """newton()'s 'Singular' branch (artery_network.py:701-708), golden vector made by the
    reference's own newton() on a state whose Jacobian LAPACK finds exactly singular."""
    from bloodflow_b200 import _ffi
    g = np.load(os.path.join(GOLDEN, 'bc_%s.npz' % case))
    an = _facade(case)
    Un = g['s3_Un']
    an._dev.set_state(Un[None, :, :, 0], Un[None, :, :, 1])
    an._dev.clear_flags()
    x, solves = an._dev.junction_newton(0, g['singular_dex'], g['singular_x_in'], k_max=1)
    assert solves == 1
    assert int(an._dev.get_flags()[0]) & _ffi.AF_FLAG_SINGULAR
    # cond(J + 1e-6 I) ~ 1e9 (two rows are 1e-6 on the diagonal and nothing else): two correct LU
    # factorisations agree to ~1e-5 of the largest entry; tests/test_kernel_math_host.py checks the
    # backward error of the same device code on the host
    want = g['singular_newton_k1']
    np.testing.assert_allclose(x, want, rtol=0, atol=1e-3 * np.abs(want).max())
    an._dev.clear_flags()
    assert int(an._dev.get_flags()[0]) == 0


def test_failure_is_raised_at_the_failing_step():
    """DOLFIN raises inside the failing step (artery.py:244).  The shipped demo cfg with HEAD's
    Windkessel factors goes non-finite in the first step: solve() must come back at once
    -- not after the 16 000 steps of the run -- and name step 0."""
    import time
    import bloodflow_b200.arteryfe as af
    p = af.ParamParser(os.path.join(ROOT, 'config', 'demo_arterybranch.cfg'))
    an = af.ArteryNetwork(p)                    # HEAD factors 0.3 / 0.01
    t0 = time.perf_counter()
    with pytest.raises(RuntimeError) as e:
        an.solve(write_output=False)
    dt = time.perf_counter() - t0
    assert an.failed_step == 0
    assert 'time step 0' in str(e.value)
    # 16 000 steps of this network take ~0.5 s; a loop that stops enqueuing and whose queued launches
    # return immediately is far below that
    total = an.geo['Nt'] * an.geo['N_cycles']
    assert total >= 4000 and dt < 0.25, dt
    # the handle is usable again after the flags are cleared
    an2 = af.ArteryNetwork(p, wk_scale=(1.0, 1.0))
    res = an2.solve(write_output=False, n_steps=40)
    assert np.isfinite(res['flow']).all()


def test_flags_are_not_sticky_across_artery_solves():
    an = _facade('c1')
    a = an.arteries[1]
    a.U_in = [-1.0, 0.2]                 # a negative area: the Newton iteration goes non-finite
    with pytest.raises(RuntimeError):
        a.solve()
    a.U_in = [a.A0(0), a.q0]
    a.solve()                            # the earlier failure was reported once; this solve is fine
    a.update_solution()


def test_root_inlet_holder_keeps_the_last_sample():
    """arteries[0].q_in = q_in (artery_network.py:777): after stepping, the root's holder has
    the inlet value of the last step."""
    an = _facade('c1')
    an.define_x()
    an._dev.step(0, 7)
    assert an.arteries[0].q_in == an.q_ins[7 % an.geo['Nt']]


PDL_WORKER = r'''
import sys, os
sys.path.insert(0, %(root)r)
os.chdir(%(root)r)
import numpy as np
import bloodflow_b200.arteryfe as af
from bloodflow_b200 import ensemble as en
from bloodflow_b200.synthetic import tree_param
out = {}
for name, p, steps in (('c1', af.ParamParser('tests/golden/golden_c1.cfg'), 120),
                       ('o3', af.ParamParser('tests/golden/golden_o3.cfg'), 120),
                       ('c4', tree_param(8, Nx=1000, Nt=8000), 40)):
    an = af.ArteryNetwork(p)
    an.define_x()
    an._dev.set_counter_log(steps)
    an._dev.step(0, steps // 2)                   # two calls: the dataflow flags are re-based per call
    an._dev.step(steps // 2, steps - steps // 2)
    A, q = an._dev.get_state()
    a, j, pi = an._dev.fetch_counter_log()
    assert int(an._dev.get_flags()[0]) & 64 == 0  # AF_FLAG_SYNC_TIMEOUT
    out.update({name + '_A': A, name + '_q': q, name + '_x': an._dev.get_junction_x(), name + '_bc': an._dev.get_bc(),
                name + '_dex': an._dev.get_dex(), name + '_a': a, name + '_j': j, name + '_p': pi})
ens = en.EnsembleNetwork(tree_param(4, Nx=100, Nt=2000), 700, first_index=5)
ens.dev.set_counter_log(30)
ens.dev.step(0, 30)
A, q = ens.dev.get_state()
a, j, pi = ens.dev.fetch_counter_log()
out.update(ens_A=A, ens_q=q, ens_x=ens.dev.get_junction_x(), ens_bc=ens.dev.get_bc(), ens_a=a, ens_j=j, ens_p=pi)
np.savez(sys.argv[1], **out)
'''


def test_plain_stream_order_equals_programmatic_dependent_launch(tmp_path):
    """The default scheduling (programmatic dependent launch + per-artery / per-junction dataflow flags
    on multi-warp single networks) against AF_NO_DATAFLOW=1 (dependent launch with grid-wide waits) and
    AF_NO_PDL=1 (plain stream order between the two kernels of a step): state, junction unknowns, Dirichlet data, dex and every counter
    bitwise equal on the demo bifurcation, the order-3 tree, a short run of the order-8 tree and a
    700-member ensemble.  Guards the 'only constants before the grid-dependency wait' rule of
    k_boundary / k_artery (af_kernels.cu)."""
    script = tmp_path / 'pdl_worker.py'
    script.write_text(PDL_WORKER % {'root': ROOT})
    outs = []
    for tag, env in (('default', {}), ('no_dataflow', {'AF_NO_DATAFLOW': '1'}), ('plain', {'AF_NO_PDL': '1'})):
        path = str(tmp_path / (tag + '.npz'))
        r = subprocess.run([sys.executable, str(script), path], env=dict(os.environ, **env), capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(np.load(path))
    for other in outs[1:]:
        assert set(outs[0].files) == set(other.files)
        for key in outs[0].files:
            np.testing.assert_array_equal(outs[0][key], other[key], err_msg=key)
