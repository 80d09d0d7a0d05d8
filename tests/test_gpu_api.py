"""The arteryfe API on the GPU beyond the golden cases: the reference's own
unit-test identities re-expressed against the HEAD API, edge
cases of the network shape, the output layout, and error behaviour of the C ABI."""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _net(case='c1', **kw):
    import bloodflow_b200.arteryfe as af
    return af.ArteryNetwork(af.ParamParser(os.path.join(GOLDEN, 'golden_%s.cfg' % case)), **kw)


def _advance(an, n):
    an.define_x()
    an._dev.step(0, n)
    an._dev.sync()


def test_problem_function_rows_vanish():
    """reference test_problem_function (test/test_artery_network.py:257-398):
    each of the 18 rows is zero once its unknown is set from the others."""
    an = _net('c1')
    _advance(an, 12)
    p, d1, d2 = an.arteries
    an.adjust_bifurcation_step(p, d1, d2)
    L = p.param['L']
    x = np.array(an.x[0]) * (1 + 0.02 * np.cos(np.arange(18)))
    Uh = [an.compute_U_half(p, L - p.dex, L, p.Un(L - p.dex), p.Un(L)),
          an.compute_U_half(d1, 0, d1.dex, d1.Un(0), d1.Un(d1.dex)),
          an.compute_U_half(d2, 0, d2.dex, d2.Un(0), d2.Un(d2.dex))]
    F = lambda z: an.problem_function(p, d1, d2, z)       # noqa: E731
    for v in range(3):
        z = x.copy(); z[3 * v + 1] = (Uh[v][1] + z[3 * v + 2]) / 2        # rows 0-2
        assert abs(F(z)[v]) < 1e-12
        z = x.copy(); z[10 + 3 * v] = (Uh[v][0] + z[11 + 3 * v]) / 2      # rows 3-5
        assert abs(F(z)[3 + v]) < 1e-12
    z = x.copy(); z[0] = z[3] + z[6]
    assert abs(F(z)[6]) < 1e-13
    z = x.copy(); z[1] = z[4] + z[7]
    assert abs(F(z)[7]) < 1e-13
    # pressure continuity rows: solve A_d from A_p
    fp, A0p = p.f(L), p.A0(L)
    for row, ip, idd, d in ((8, 10, 13, d1), (9, 10, 16, d2), (10, 9, 12, d1), (11, 9, 15, d2)):
        z = x.copy()
        pp = fp * (1 - np.sqrt(A0p / z[ip]))
        z[idd] = d.A0(0) / (1 - pp / d.f(0)) ** 2
        assert abs(F(z)[row]) < 1e-8 * fp
    # Lax-Wendroff rows: x[3v] and x[9+3v] enter linearly with coefficient 1
    y = F(x)
    for v in range(3):
        z = x.copy(); z[3 * v] -= y[12 + v]
        assert abs(F(z)[12 + v]) < 1e-12
        z = x.copy(); z[9 + 3 * v] -= y[15 + v]
        assert abs(F(z)[15 + v]) < 1e-12


def test_jacobian_against_finite_differences():
    """reference test_jacobian (:401-431) at x = ones(18) (where A = 1 hides the
    missing /A of the flux) and, to document SURVEY F6, at a physical x where the
    hand-coded Jacobian of q^2/A does NOT match the q^2 flux of problem_function."""
    an = _net('c1')
    _advance(an, 12)
    p, d1, d2 = an.arteries
    an.adjust_bifurcation_step(p, d1, d2)
    h = 1e-7

    def fd(x):
        J = np.zeros((18, 18))
        f0 = an.problem_function(p, d1, d2, x)
        for j in range(18):
            e = np.zeros(18); e[j] = h
            J[:, j] = (an.problem_function(p, d1, d2, x + e) - f0) / h
        return J

    x = np.ones(18)
    J, Jfd = an.jacobian(p, d1, d2, x), fd(x)
    scale = np.maximum(np.abs(J), 1.0)
    # the Jacobian evaluates the Expressions at L+dex / -dex, the function at +-dex/2 (quirk 3):
    # on the tapered daughters that is a ~1e-3 relative difference, the reference's own tolerance
    assert np.max(np.abs(J - Jfd) / scale) < 2e-3
    x = np.array(an.x[0])
    J, Jfd = an.jacobian(p, d1, d2, x), fd(x)
    scale = np.maximum(np.abs(J), 1.0)
    err = np.abs(J - Jfd) / scale
    bad = {(12, 2), (12, 11), (13, 5), (13, 14), (14, 8), (14, 17)}       # the q^2 vs q^2/A entries
    for i in range(18):
        for j in range(18):
            if (i, j) not in bad:
                assert err[i, j] < 2e-3, (i, j, J[i, j], Jfd[i, j])
    assert max(err[i, j] for i, j in bad) > 1e-2                          # F6 is real


def test_newton_fixed_point_and_bc_holders():
    """reference test_newton (:434-455) and test_set_inner_bc (:458-509)."""
    an = _net('o3')
    _advance(an, 20)
    for k, ip in enumerate(an.range_parent_arteries):
        i1, i2 = an.daughter_arteries(ip)
        p, d1, d2 = an.arteries[ip], an.arteries[i1], an.arteries[i2]
        an.adjust_bifurcation_step(p, d1, d2)
        # CFL holds at the three ends
        L = p.param['L']
        assert p.check_CFL(L, *p.Un(L)) and d1.check_CFL(0, *d1.Un(0)) and d2.check_CFL(0, *d2.Un(0))
        x = an.newton(p, d1, d2, an.initial_x(p, d1, d2), 1000, 1e-14)
        FF = an.problem_function(p, d1, d2, x)
        assert np.all(np.abs(FF) < 1e-9) and np.all(x > 1e-12)
        an.set_inner_bc(ip, i1, i2)
        xk = np.array(an.x[k])
        np.testing.assert_array_equal(p.U_out, [xk[9], xk[0]])
        np.testing.assert_array_equal(d1.U_in, [xk[12], xk[3]])
        np.testing.assert_array_equal(d2.U_in, [xk[15], xk[6]])


def test_windkessel_changes_outlet_pressure_mildly():
    """reference test_compute_A_out (:197-213): one call moves the outlet
    pressure by less than 10 %; the pre-HEAD name is kept as an alias."""
    an = _net('c2')
    _advance(an, 30)
    for i in an.range_leaf_arteries:
        a = an.arteries[i]
        p_old = a.compute_outlet_pressure(a.Un(a.param['L'])[0])
        A_out = an.compute_A_out(a)
        assert abs(a.compute_outlet_pressure(A_out) / p_old - 1) < 0.1
        assert 1 <= an.last_picard_its <= 100


def test_adjust_dex_margins():
    """reference test_adjust_dex (test/test_artery.py:174-212)."""
    an = _net('c1')
    a = an.arteries[1]
    x, (A, q) = 0.3 * a.param['L'], a.Un(0.3 * a.param['L'])
    M = a.CFL_term(x, A, q)
    for margin in (0.1, 0.05, 1e-4, 1e-8):
        a.adjust_dex(x, A, q, margin)
        np.testing.assert_allclose(a.dex, (1 + margin) * a.dt / M, rtol=1e-14)
        assert a.check_CFL(x, A, q)
    a.dex = 0.99 * a.dt / M
    assert not a.check_CFL(x, A, q)


def test_standalone_artery_like_the_reference_unit_tests():
    """test/test_artery.py: an Artery built on its own, one solve, Dirichlet values hold."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200.arteryfe.artery import Artery
    from bloodflow_b200.arteryfe.utils import nondimensionalise_parameters, is_near
    p = af.ParamParser(os.path.join(GOLDEN, 'golden_c1.cfg'))
    nd = nondimensionalise_parameters(p)
    T = 0.917 * nd['qc'] / nd['rc'] ** 3
    for root, leaf in ((True, False), (False, True), (False, False)):
        a = Artery(1, T, nd, root=root, leaf=leaf)
        a.define_geometry(p.geo)
        a.define_solution(0.2, 0.55)
        L = a.param['L']
        for x in a.dx * np.array([0, 3, 17, a.Nx]):       # Un is the P1 interpolant: exact at the vertices
            assert is_near(a.Un(x)[0], a.A0(x)) and is_near(a.Un(x)[1], 0.2)
            assert is_near(a.pn(x), a.param['p0'])
        if root:
            a.q_in = 0.21
        else:
            a.U_in = [a.A0(0) * 1.01, 0.21]
        if leaf:
            a.A_out = a.A0(L) * 1.005
        else:
            a.U_out = [a.A0(L) * 1.005, 0.2]
        a.solve()
        if root:
            assert is_near(a.U(0)[1], a.q_in)
        else:
            assert is_near(a.U(0)[0], a.U_in[0]) and is_near(a.U(0)[1], a.U_in[1])
        assert is_near(a.U(L)[0], a.A_out if leaf else a.U_out[0])
        assert is_near(a.Un(0.5 * L)[1], 0.2)           # Un untouched until update_solution
        a.update_solution()
        np.testing.assert_array_equal(a.Un.values(), a.U.values())
        a.update_pressure()
        A = a.Un.values()[:, 0]
        xs = a.dx * np.arange(a.Nx + 1)
        np.testing.assert_allclose(a.pn.values(), a.param['p0'] + a.f(xs) * (1 - np.sqrt(a.A0(xs) / A)), rtol=1e-12)


@pytest.mark.parametrize('Nx', (2, 3, 5, 31, 127, 128, 129, 513, 1500, 2047))     # 1500, 2047: the 512-thread CTA
def test_vertex_counts_around_the_thread_mapping(Nx):
    """Four vertices per thread, T a power of two: every padding situation
    against the C oracle for 6 steps on an order-3 tree."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200.synthetic import tree_param
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle
    p = tree_param(3, Nx=Nx, Nt=6400)
    an = af.ArteryNetwork(p)
    an.define_x()
    an._dev.set_counter_log(6)
    an._dev.step(0, 6)
    A, q = an._dev.get_state()
    p = tree_param(3, Nx=Nx, Nt=6400)
    c = COracle(ao.OracleNetwork(p.param, p.geo, p.solution))
    _, la, lj = c.step(0, 6, nthreads=1, log=True)
    Ao, qo = c.state()
    np.testing.assert_allclose(A[0], Ao, rtol=1e-10)
    np.testing.assert_allclose(q[0], qo, rtol=1e-9, atol=1e-12)
    art, jun, _ = an._dev.fetch_counter_log()
    np.testing.assert_array_equal(art[:, 0], la)
    np.testing.assert_array_equal(jun[:, 0], lj)


def test_incomplete_tree():
    """Ru == 0 marks a missing vessel (artery_network.py:90-91): order 3 with the
    second daughter pair absent, so artery 2 is a leaf one level up."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200.synthetic import tree_param
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle

    def make():
        p = tree_param(3, Nx=48, Nt=1600)
        p.param['Ru'][5:] = 0.0
        p.param['Rd'][5:] = 0.0
        p.param['Rd'][2] = 0.37 * 0.8 / (0.37 * 0.8)      # stays the multiplier 1 -> untapered via alpha rule
        return p
    an = af.ArteryNetwork(make())
    assert an.arteries[5] is None and an.arteries[6] is None
    assert an.range_parent_arteries == [0, 1] and an.range_leaf_arteries == [2, 3, 4]
    assert an._dev.na == 5 and an._dev.nj == 2 and an._dev.nl == 3
    an.define_x()
    an._dev.step(0, 15)
    A, q = an._dev.get_state()
    pp = make()
    c = COracle(ao.OracleNetwork(pp.param, pp.geo, pp.solution))
    c.step(0, 15, nthreads=1)
    Ao, qo = c.state()
    np.testing.assert_allclose(A[0], Ao, rtol=1e-10)
    np.testing.assert_allclose(q[0], qo, rtol=1e-9, atol=1e-12)


def test_output_layout_round_trip(tmp_path):
    """solve() writes data.cfg + <out>/{flow,area,pressure}/name_i.xdmf; read_output and
    XDMF_to_matrix (utils.py:203-272) give back M[Nx+1, Nt] (reference test_dump_metadata)."""
    import bloodflow_b200.arteryfe as af
    p = af.ParamParser(os.path.join(GOLDEN, 'golden_c1.cfg'))
    out = str(tmp_path / 'run')
    p.solution['output_location'] = out
    an = af.ArteryNetwork(p)
    res = an.solve()
    order, Nx, Nt, T0, T, L, rc, qc, rho, mesh_locations, names, locations = af.read_output(out + '/data.cfg')
    assert (order, Nx, Nt) == (2, 40, 40) and names == ['flow', 'area', 'pressure']
    assert T0 == 0.0 and abs(T - an.T) < 1e-12 and len(L) == 3 and len(mesh_locations) == 3
    np.testing.assert_allclose(L, an.nondim['L'])
    for name, loc in zip(names, locations):
        for i in range(3):
            M = af.XDMF_to_matrix(Nx, Nt, mesh_locations[i], '%s/%s_%i.xdmf' % (loc, name, i), name)
            np.testing.assert_array_equal(M, res[name][:, i, :].T)
    np.testing.assert_array_equal(np.load(mesh_locations[1]), an.arteries[1].dx * np.arange(Nx + 1))
    assert '<Xdmf' in open(out + '/flow/flow_0.xdmf').read()


def test_state_and_holder_round_trips():
    an = _net('o3')
    dev = an._dev
    A, q = dev.get_state()
    dev.set_state(A * 1.01, q * 0.5)
    A2, q2 = dev.get_state()
    np.testing.assert_array_equal(A2, A * 1.01)
    np.testing.assert_array_equal(q2, q * 0.5)
    bc = dev.get_bc(); bc[0, 3, :] = [1.0, 2.0, 3.0, 4.0]; dev.set_bc(bc)
    np.testing.assert_array_equal(dev.get_bc(), bc)
    a = an.arteries[3]
    assert list(a.U_in) == [1.0, 2.0] and a.A_out == 3.0
    a.dex = 0.123
    assert a.dex == 0.123
    x = an.x
    x[1] = np.arange(18.0)
    np.testing.assert_array_equal(np.array(an.x[1]), np.arange(18.0))


def test_abi_error_codes():
    from bloodflow_b200 import _ffi
    lib = _ffi.lib
    an = _net('c1')
    h = an._dev._h
    out = (ctypes.c_double * 4)()
    x = (ctypes.c_double * 1)(0.0)
    assert lib.af_eval_state(h, 0, 99, 0, 1, x, out) == _ffi.AF_EINVAL
    assert b'out of range' in lib.af_last_error()
    assert lib.af_net_step(h, -1, 1) == _ffi.AF_EINVAL
    assert lib.af_phase_arteries(h, 10 ** 6) == _ffi.AF_EINVAL
    assert lib.af_net_sync(None) == _ffi.AF_EINVAL
    d = _ffi.AfDesc()
    lib.af_desc_defaults(ctypes.byref(d))
    hh = _ffi.Handle()
    assert lib.af_net_create(ctypes.byref(d), ctypes.byref(hh)) == _ffi.AF_EINVAL     # empty descriptor
    d.abi_version = 99
    assert lib.af_net_create(ctypes.byref(d), ctypes.byref(hh)) == _ffi.AF_EINVAL
    # a pending single-artery solve blocks the fused time loop until it is assigned
    an.arteries[0].solve()
    assert lib.af_net_step(h, 0, 1) == _ffi.AF_ESTATE
    an.arteries[0].update_solution()
    assert lib.af_net_step(h, 0, 1) == _ffi.AF_OK


def test_nan_is_flagged_not_hidden():
    """Non-finite state: the artery Newton stops with the NOCONV|NONFINITE bits and
    solve() raises, like DOLFIN's RuntimeError; the reference's BC code would
    propagate NaN silently."""
    an = _net('c1')
    A, q = an._dev.get_state()
    A[0, 1, 5] = np.nan
    an._dev.set_state(A, q)
    an._dev.step(0, 1)
    an._dev.sync()
    assert int(an._dev.get_flags()[0]) & 9 == 9
    with pytest.raises(RuntimeError):
        an.arteries[1].solve()


def test_fused_step_equals_two_kernel_step(monkeypatch):
    """AF_FUSE=1 makes af_net_step run set_bcs inside the artery CTAs (one launch
    per step, every junction solved redundantly by the CTAs of its three
    vessels); the default two-kernel path must give bitwise the same states,
    junction unknowns, holders and counters."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200.synthetic import tree_param

    def run(no_fuse):
        monkeypatch.setenv('AF_FUSE', '0' if no_fuse else '1')
        an = af.ArteryNetwork(tree_param(4, Nx=300, Nt=4000))
        an.define_x()
        an._dev.set_counter_log(30)
        an._dev.step(0, 30)
        return (an._dev.get_state(), an._dev.get_junction_x(), an._dev.get_bc(), an._dev.get_dex(),
                an._dev.fetch_counter_log(), an._dev.get_totals())

    fused, plain = run(False), run(True)
    for a, b in zip(fused[0], plain[0]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(fused[1], plain[1])
    np.testing.assert_array_equal(fused[2][0, 1:], plain[2][0, 1:])       # the root's q_in holder is only
    np.testing.assert_array_equal(fused[2][0, 0, 2:], plain[2][0, 0, 2:])  # written by the fused path
    np.testing.assert_array_equal(fused[3], plain[3])
    for a, b in zip(fused[4], plain[4]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(fused[5], plain[5])


def test_streamed_store_equals_plain_fetch():
    """af_net_set_store_host streams snapshots to host buffers while af_net_step
    is still enqueueing (pinned ring of 4 slots, more snapshots than slots, the
    loop issued in several calls); the result must equal the plain fetch."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200 import _ffi
    from bloodflow_b200.synthetic import tree_param

    def run(stream):
        an = af.ArteryNetwork(tree_param(3, Nx=64, Nt=400, Nt_store=50))
        an.define_x()
        dev = an._dev
        fields = _ffi.AF_STORE_FLOW | _ffi.AF_STORE_PRESSURE | _ffi.AF_STORE_OUTLET_PRESSURE
        dev.set_store(an.store_mask(), 0, 23, fields)
        if stream:
            dev.stream_store_to_host()
        for first, n in ((0, 7), (7, 120), (127, 53)):
            dev.step(first, n)
        return dev.fetch_snapshots()

    a, b = run(True), run(False)
    assert len(a['step']) == 23 and list(a['step']) == list(b['step']) == [8 * k for k in range(23)]
    for key in ('flow', 'pressure', 'outlet_pressure'):
        np.testing.assert_array_equal(a[key], b[key])
    assert 'area' not in a


def _raw_network(junctions, leaves, na=3, Nx=16, root_artery=0, **kw):
    from bloodflow_b200.engine import DeviceNetwork
    ones = np.ones(na)
    return DeviceNetwork(Nx, 8, 1.0, 1e-3, junctions, leaves, k1=2e7, k2=-22.5, k3=8.65e5, rho=1.0, Re=100.0,
                         db=1.0, p0=1.0, Ru=0.4 * ones, Rd=0.4 * ones, L=5 * ones, q0=ones,
                         R1=np.ones(len(leaves)), R2=np.ones(len(leaves)), CT=np.ones(len(leaves)),
                         q_ins=np.ones(8), root_artery=root_artery, **kw)


@pytest.mark.parametrize('junctions,leaves,root', [
    ([[0, 1, 1]], [1, 2], 0),                 # the same daughter twice
    ([[0, 1, 2], [0, 1, 2]], [1, 2], 0),      # two junctions on the same ends
    ([[0, 1, 2]], [0, 1, 2], 0),              # a parent that is also an outlet
    ([[0, 1, 2]], [1, 1], 0),                 # an outlet listed twice
    ([[1, 0, 2]], [0, 2], 0),                 # the root as a daughter
    ([[0, 1, 2]], [1, 2], 3),                 # root index outside the network
    ([[0, 1, 7]], [1, 2], 0),                 # index out of range
])
def test_inconsistent_topologies_are_rejected(junctions, leaves, root):
    from bloodflow_b200 import _ffi
    with pytest.raises(_ffi.AfError) as e:
        _raw_network(junctions, leaves, root_artery=root)
    assert e.value.code == _ffi.AF_EINVAL


def test_size_limits():
    from bloodflow_b200 import _ffi
    net = _raw_network([[0, 1, 2]], [1, 2], Nx=2047)      # 2048 vertices: the largest CTA (512 threads)
    net.step(0, 2)
    net.sync()
    A, q = net.get_state()
    assert np.isfinite(A).all() and np.isfinite(q).all()
    net.close()
    with pytest.raises(_ffi.AfError):
        _raw_network([[0, 1, 2]], [1, 2], Nx=2048)
    with pytest.raises(_ffi.AfError):
        _raw_network([[0, 1, 2]], [1, 2], Nx=1)


def test_store_reconfiguration_does_not_accumulate_device_memory():
    import torch
    import bloodflow_b200
    from bloodflow_b200 import _ffi
    net = _raw_network([[0, 1, 2]], [1, 2], Nx=1023)
    mask = np.ones(8, dtype=np.uint8)
    fields = _ffi.AF_STORE_FLOW | _ffi.AF_STORE_AREA | _ffi.AF_STORE_PRESSURE
    net.set_store(mask, 0, 4096, fields)                  # 3 x 4096 x 3 x 1024 x 8 B = 302 MB
    net.step(0, 3)
    first = net.fetch_snapshots()
    free0 = torch.cuda.mem_get_info(0)[0]
    for cap in (4097, 4098, 4099, 4100):
        net.set_store(mask, 0, cap, fields)
    bloodflow_b200.release_cached_memory()                # released blocks sit in the library's cache
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 64 << 20                       # the old buffers were released, not kept
    net.step(3, 2)
    again = net.fetch_snapshots()
    assert list(first['step']) == [0, 1, 2] and list(again['step']) == [3, 4]
    net.close()


def test_recycled_blocks_do_not_leak_state_between_networks():
    """Blocks of a destroyed handle are reused by the next same-shaped one
    (af_release_cached_memory): a network built on recycled memory must give
    bitwise the results of one built on fresh memory."""
    import bloodflow_b200

    def run(scale):
        net = _raw_network([[0, 1, 2]], [1, 2], Nx=200)
        if scale != 1.0:
            A, q = net.get_state()
            net.set_state(A * scale, q * scale)
        net.set_store(np.ones(8, dtype=np.uint8), 0, 16, 1 | 2 | 4 | 8)
        net.step(0, 12)
        out = (net.get_state(), net.get_bc(), net.get_dex(), net.get_junction_x(), net.fetch_snapshots())
        net.close()
        return out

    bloodflow_b200.release_cached_memory()
    fresh = run(1.0)
    run(1.3)                                   # leaves different data in the cached blocks
    recycled = run(1.0)
    for a, b in zip(fresh[0], recycled[0]):
        np.testing.assert_array_equal(a, b)
    for k in (1, 2, 3):
        np.testing.assert_array_equal(fresh[k], recycled[k])
    for key in ('flow', 'area', 'pressure', 'outlet_pressure'):
        np.testing.assert_array_equal(fresh[4][key], recycled[4][key])
    bloodflow_b200.release_cached_memory()


@pytest.mark.parametrize('which,ref', [(0, lambda x: 1.0 / x), (1, lambda x: 1.0 / np.sqrt(x))])
def test_branch_free_reciprocals_are_accurate_to_an_ulp(which, ref):
    """af_rcp_fast / af_rsqrt_fast (hardware seed + one cubic step) against the
    IEEE expressions over 24 decades: < 1 ulp (relative 2.2e-16)."""
    from bloodflow_b200 import _ffi
    rng = np.random.default_rng(7)
    x = np.ascontiguousarray(10.0 ** rng.uniform(-12, 12, 200000) * rng.uniform(1.0, 2.0, 200000))
    out = np.empty_like(x)
    _ffi.check(_ffi.lib.af_eval_math(0, which, len(x), _ffi.dptr(x), _ffi.dptr(out)))
    err = np.abs(out / ref(x) - 1.0)
    assert err.max() < 2.3e-16, err.max()
    assert np.mean(out == ref(x)) > 0.5            # most results are the correctly rounded value


def test_cache_can_be_disabled(tmp_path):
    """AF_CACHE_MB=0: released blocks go straight back to the driver (checked in a
    fresh process, the cap is read once per process)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, torch\n"
        "from bloodflow_b200.engine import DeviceNetwork\n"
        "o = np.ones(3)\n"
        "net = DeviceNetwork(1023, 8, 1.0, 1e-3, [[0, 1, 2]], [1, 2], k1=2e7, k2=-22.5, k3=8.65e5, rho=1.0, Re=100.0,\n"
        "                    db=1.0, p0=1.0, Ru=0.4 * o, Rd=0.4 * o, L=5 * o, q0=o, R1=np.ones(2), R2=np.ones(2),\n"
        "                    CT=np.ones(2), q_ins=np.ones(8))\n"
        "net.set_store(np.ones(8, dtype=np.uint8), 0, 2048, 7)\n"
        "net.step(0, 4); net.sync()\n"
        "used = torch.cuda.mem_get_info(0)[0]\n"
        "net.close()\n"
        "freed = torch.cuda.mem_get_info(0)[0] - used\n"
        "assert freed > 100 << 20, freed\n"
        "print('freed', freed >> 20, 'MB')\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, AF_CACHE_MB='0')
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


def test_schedule_follows_the_shape_of_the_network():
    """af_net_schedule: a multi-warp single network runs the persistent kernel pair, a one-warp network
    (41 vertices) and an ensemble the per-step launches with grid-wide waits."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200 import _ffi
    from bloodflow_b200 import ensemble as en
    from bloodflow_b200.synthetic import tree_param
    tree = af.ArteryNetwork(tree_param(4, Nx=400, Nt=2000))
    assert tree._dev.schedule() == _ffi.AF_SCHEDULE_PERSISTENT
    small = af.ArteryNetwork(af.ParamParser(os.path.join(GOLDEN, 'golden_c1.cfg')))
    assert small._dev.schedule() == _ffi.AF_SCHEDULE_PER_STEP
    ens = en.EnsembleNetwork(tree_param(4, Nx=100, Nt=2000), 8)
    assert ens.dev.schedule() == _ffi.AF_SCHEDULE_PER_STEP
