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
for name, p, steps in (('c1', af.ParamParser('tests/golden/golden_c1.cfg'), 200),      # > one 64-step chunk per call
                       ('o3', af.ParamParser('tests/golden/golden_o3.cfg'), 200),
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
    """The default scheduling of multi-warp single networks (persistent kernels over chunks of steps,
    per-artery flags, Dirichlet mailbox, leaves' Windkessel inside their artery CTA) against AF_NO_PERSIST=1
    (one launch pair per step with the same flags and mailbox), AF_NO_DATAFLOW=1 (dependent launch with
    grid-wide waits) and AF_NO_PDL=1 (plain stream order between the two kernels of a step): state,
    junction unknowns, Dirichlet data, dex and every counter
    bitwise equal on the demo bifurcation, the order-3 tree, a short run of the order-8 tree and a
    700-member ensemble.  Guards the 'only constants before the grid-dependency wait' rule of
    k_boundary / k_artery (af_kernels.cu)."""
    script = tmp_path / 'pdl_worker.py'
    script.write_text(PDL_WORKER % {'root': ROOT})
    outs = []
    for tag, env in (('default', {}), ('no_persist', {'AF_NO_PERSIST': '1'}), ('no_dataflow', {'AF_NO_DATAFLOW': '1'}),
                     ('plain', {'AF_NO_PDL': '1'})):
        path = str(tmp_path / (tag + '.npz'))
        r = subprocess.run([sys.executable, str(script), path], env=dict(os.environ, **env), capture_output=True,
                           text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append(np.load(path))
    for other in outs[1:]:
        assert set(outs[0].files) == set(other.files)
        for key in outs[0].files:
            np.testing.assert_array_equal(outs[0][key], other[key], err_msg=key)


def _compare_with_c_oracle(dev, c, total, chunk, counts_exact=True):
    """Step the device network and the C oracle side by side; state and junction unknowns to 1e-9 at
    every checkpoint, artery Newton counts equal at every step, junction counts up to near-ties."""
    dev.set_counter_log(total)
    la, lj, done = [], [], 0
    while done < total:
        n = min(chunk, total - done)
        dev.step(done, n)
        failed, a_, j_ = c.step(done, n, nthreads=0, log=True)
        assert failed == 0
        la.append(a_)
        lj.append(j_)
        done += n
        A, q = dev.get_state()
        Ao, qo = c.state()
        np.testing.assert_allclose(A[0], Ao, rtol=1e-9, err_msg='A at step %d' % done)
        np.testing.assert_allclose(q[0], qo, rtol=1e-9, atol=1e-9 * np.abs(qo).max(), err_msg='q at step %d' % done)
        np.testing.assert_allclose(dev.get_junction_x()[0], c.x(), rtol=1e-9)
    art, jun, _ = dev.fetch_counter_log()
    np.testing.assert_array_equal(art[:, 0], np.concatenate(la))
    lj = np.concatenate(lj)
    mism = np.count_nonzero(jun[:, 0] != lj)
    if counts_exact:
        assert mism == 0, mism
    else:
        assert mism <= 1e-3 * lj.size and np.abs(jun[:, 0] - lj).max() <= 1, mism
    assert int(dev.get_flags()[0]) & (1 | 8) == 0
    return mism


def test_shipped_configs_full_horizon():
    """BASELINE configs C1 and C2 at their stated size AND horizon against the C oracle: the shipped
    demo_arterybranch.cfg (3 x 401 vertices) over its whole cardiac cycle of 4 000 steps with the documented
    Windkessel factors (1, 1) (HEAD's factors diverge in step 0), and demo_arterybranch_geo.cfg
    (+ R_term = 0.296, HEAD factors; 3 x 201 vertices) over all of its 4 cycles = 8 000 steps.  Neither
    amplifies round-off (a 1e-13 perturbation stays below 4e-12 over the horizon, C oracle against itself)."""
    import bloodflow_b200.arteryfe as af
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle
    cfg = 'config/demo_arterybranch.cfg'
    an = af.ArteryNetwork(af.ParamParser(cfg), wk_scale=(1, 1))
    an.define_x()
    assert (an._dev.na, an._dev.np, an.geo['Nt']) == (3, 401, 4000)
    c = COracle(ao.network_from_cfg(cfg, wk_scale=(1, 1)))
    _compare_with_c_oracle(an._dev, c, 4000, 500, counts_exact=False)
    c.close()

    cfg = 'config/demo_arterybranch_geo.cfg'
    p = af.ParamParser(cfg)
    p.param['R_term'] = 0.296
    an = af.ArteryNetwork(p)
    an.define_x()
    assert (an._dev.na, an._dev.np, an.geo['Nt'], an.geo['N_cycles']) == (3, 201, 2000, 4)
    c = COracle(ao.network_from_cfg(cfg, R_term=0.296))
    _compare_with_c_oracle(an._dev, c, 8000, 1000, counts_exact=False)
    c.close()


def test_c5_members_over_both_cycles_match_c_oracle():
    """BASELINE config 5 at its stated horizon: members of a batch of perturbed order-4 networks (Nx=100,
    Nt=2000) over BOTH cardiac cycles (4 000 steps) against the C oracle run on the same perturbed
    parameters -- state at the end, the stored outlet pressures of every leaf at all 200 dump steps, and
    the artery Newton counts of every step."""
    from bloodflow_b200 import ensemble as en
    from bloodflow_b200.synthetic import tree_param
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle
    base = tree_param(4, Nx=100, Nt=2000, N_cycles=2, Nt_store=100)
    first, M, steps = 30000, 640, 4000
    ens = en.EnsembleNetwork(base, M, first_index=first)
    dev = ens.dev
    ens.configure_store(2, 2)                       # both cycles: 200 snapshots
    dev.set_counter_log(steps)
    dev.step(0, steps)
    A, q = dev.get_state()
    art, jun, _ = dev.fetch_counter_log()
    snaps = dev.fetch_snapshots()
    op = snaps['outlet_pressure']                   # [snapshot, member, leaf]
    leaves = np.asarray(ens.description['leaves'])
    mask = ens.nominal.store_mask()
    dump_steps = np.flatnonzero(np.tile(mask, 2))
    assert op.shape[0] == len(dump_steps) == 200
    np.testing.assert_array_equal(snaps['step'], dump_steps)
    for m in (0, M // 3, M - 1):
        p = en.perturb_param(base, first + m)
        c = COracle(ao.OracleNetwork(p.param, p.geo, p.solution))
        la, lj, po = [], [], []
        done = 0
        for s in dump_steps:
            # the reference stores Un of time t_s, i.e. BEFORE the solve of step s (an:920-938)
            if s > done:
                failed, a_, j_ = c.step(done, s - done, nthreads=0, log=True)
                assert not failed
                la.append(a_)
                lj.append(j_)
                done = s
            po.append(c.pressure()[leaves, -1])
        failed, a_, j_ = c.step(done, steps - done, nthreads=0, log=True)
        la.append(a_)
        lj.append(j_)
        Ao, qo = c.state()
        np.testing.assert_allclose(A[m], Ao, rtol=1e-9, err_msg='member %d' % m)
        np.testing.assert_allclose(q[m], qo, rtol=1e-9, atol=1e-9 * np.abs(qo).max(), err_msg='member %d' % m)
        np.testing.assert_allclose(op[:, m], np.array(po), rtol=1e-9, err_msg='outlet pressure, member %d' % m)
        np.testing.assert_array_equal(art[:, m], np.concatenate(la))
        lj = np.concatenate(lj)
        assert np.count_nonzero(jun[:, m] != lj) <= 1e-3 * lj.size and np.abs(jun[:, m] - lj).max() <= 1
        c.close()
    assert ens.flagged_members() == 0


def test_two_networks_stepped_concurrently_share_the_device():
    """The persistent kernel pair of a single network needs all of its CTAs resident; two pairs at once could
    starve each other.  af_net_step therefore takes a per-device lease, and a call that finds it taken uses the
    per-step launches (same results).  Two order-6 trees stepped from two threads at the same time: both
    finish without a wait time-out and equal a network stepped alone, bit for bit."""
    import threading
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200 import _ffi
    from bloodflow_b200.synthetic import tree_param

    def make():
        an = af.ArteryNetwork(tree_param(6, Nx=500, Nt=4000))
        an.define_x()
        return an

    alone = make()
    assert alone._dev.schedule() == _ffi.AF_SCHEDULE_PERSISTENT
    alone._dev.step(0, 600)
    A0, q0 = alone._dev.get_state()
    nets = [make(), make()]
    threads = [threading.Thread(target=lambda a=a: [a._dev.step(k * 100, 100) for k in range(6)]) for a in nets]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for an in nets:
        A, q = an._dev.get_state()
        assert int(an._dev.get_flags()[0]) & (1 | 8 | 64) == 0          # no failure, no AF_FLAG_SYNC_TIMEOUT
        np.testing.assert_array_equal(A, A0)
        np.testing.assert_array_equal(q, q0)
