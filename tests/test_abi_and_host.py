"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/arteryfe_b200.h declares, the ctypes struct matches the C
struct, the product fails loudly without a device, and the host-side set-up of
the facade equals the reference's (golden) values."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    from bloodflow_b200 import _ffi
    header = open(os.path.join(ROOT, 'include', 'arteryfe_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(af_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_ffi.PROTOTYPES), declared ^ set(_ffi.PROTOTYPES)
    out = subprocess.check_output(['nm', '-D', '--defined-only', _ffi.LIB_PATH]).decode()
    exported = set(re.findall(r' T (af_[a-z0-9_]+)', out))
    assert declared <= exported, declared - exported
    assert _ffi.lib.af_version() == _ffi.AF_ABI_VERSION


def test_desc_struct_layout_matches_c(tmp_path):
    """sizeof/offsetof of af_desc as gcc sees the header == the ctypes mirror."""
    from bloodflow_b200 import _ffi
    fields = [f[0] for f in _ffi.AfDesc._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "arteryfe_b200.h"\nint main(){printf("%zu", sizeof(af_desc));\n'
    for f in fields:
        src += 'printf(" %%zu", offsetof(af_desc, %s));\n' % f
    src += 'return 0;}\n'
    c = tmp_path / 'layout.c'
    c.write_text(src)
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(c), '-o', str(exe)])
    nums = [int(t) for t in subprocess.check_output([str(exe)]).decode().split()]
    import ctypes
    assert nums[0] == ctypes.sizeof(_ffi.AfDesc)
    assert nums[1:] == [getattr(_ffi.AfDesc, f).offset for f in fields]


def test_no_cpu_fallback():
    """Without a CUDA device network creation fails with AF_ENODEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200._ffi import AfError
    with pytest.raises(AfError) as e:
        af.ArteryNetwork(af.ParamParser(os.path.join(GOLDEN, 'golden_c1.cfg')))
    assert e.value.code == -2


def test_missing_library_fails_loudly(tmp_path):
    """No built CUDA library => the package refuses to import its compute layer
    (there is nothing to fall back to)."""
    import sys
    env = dict(os.environ, ARTERYFE_B200_LIB=str(tmp_path / 'nope.so'), PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, '-c', 'import bloodflow_b200.arteryfe'], env=env, capture_output=True,
                         text=True)
    assert out.returncode != 0
    assert 'ImportError' in out.stderr and 'no CPU fallback' in out.stderr.replace('There is no CPU fallback', 'no CPU fallback')


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'bloodflow_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f
                assert 'oracle/' not in text.replace('the oracle/', ''), f


@pytest.mark.parametrize('case', ('c1', 'c2', 'o3'))
def test_host_setup_matches_reference(case, monkeypatch):
    """The facade's host set-up (ParamParser, build_geometry, nondimensional
    parameters, inlet table, topology lists, q0 split) against the reference's
    own values, without touching the device."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200.arteryfe import artery_network as mod
    g = np.load(os.path.join(GOLDEN, 'setup_%s.npz' % case))
    captured = {}

    class FakeDevice(object):
        def __init__(self, *a, **k):
            captured['args'], captured['kw'] = a, k

        def set_junction_x(self, x):
            captured['x'] = x

        def close(self):
            pass

    monkeypatch.setattr(mod, 'DeviceNetwork', FakeDevice)
    an = af.ArteryNetwork(af.ParamParser(os.path.join(GOLDEN, 'golden_%s.cfg' % case)))
    for k in ('Ru', 'Rd', 'L', 'k1', 'k2', 'k3', 'Re', 'nu', 'p0', 'R1', 'R2', 'CT', 'R_term', 'p_term'):
        np.testing.assert_array_equal(np.asarray(an.nondim[k], dtype=float), g['nondim_' + k])
    assert an.T == g['T'] and an.dt == g['dt']
    np.testing.assert_array_equal(an.q_ins, g['q_ins'])
    np.testing.assert_array_equal(an.range_parent_arteries, g['range_parent'])
    np.testing.assert_array_equal(an.range_daughter_arteries, g['range_daughter'])
    np.testing.assert_array_equal(an.range_leaf_arteries, g['range_leaf'])
    arts = [a for a in an.arteries if a is not None]
    np.testing.assert_array_equal([a.dx for a in arts], g['dx'])
    np.testing.assert_array_equal([a.db for a in arts], g['db'])
    np.testing.assert_allclose([a.q0 for a in arts], g['q0'], rtol=1e-15)
    kw = captured['kw']
    np.testing.assert_allclose(np.ravel(kw['R1']), g['leaf_R1'] * 0.3, rtol=1e-15)
    np.testing.assert_allclose(np.ravel(kw['R2']), g['leaf_R2'] * 0.01, rtol=1e-15)
    np.testing.assert_array_equal(np.ravel(kw['CT']), g['leaf_CT'])
    for name in ('r0', 'A0', 'f', 'dfdr', 'drdx'):
        got = np.array([[getattr(a, name)(s * a.param['L']) for s in g['expr_frac']] for a in arts])
        np.testing.assert_allclose(got, g['expr_' + name], rtol=2e-15, atol=0)
    an.define_x()
    np.testing.assert_allclose(captured['x'][0], g['x0'], rtol=2e-15)
    # dump rule + accumulated time (artery_network.py:926, 946) against the reference loop
    run = np.load(os.path.join(GOLDEN, 'run_%s.npz' % case))
    steps = an.store_mask().nonzero()[0]
    t_all = np.concatenate([[0.0], np.cumsum(np.full(an.geo['Nt'], an.dt))])
    np.testing.assert_array_equal(t_all[steps], run['t'])


def test_shipped_configs_load(monkeypatch):
    """Both configs of the reference load: the demo verbatim, the
    _geo one needs an explicit R_term and says so."""
    import bloodflow_b200.arteryfe as af
    from bloodflow_b200.arteryfe import artery_network as mod

    class FakeDevice(object):
        def __init__(self, *a, **k):
            pass

        def set_junction_x(self, x):
            pass

    monkeypatch.setattr(mod, 'DeviceNetwork', FakeDevice)
    an = af.ArteryNetwork(af.ParamParser('config/demo_arterybranch.cfg'))
    assert an.N == 3 and an.range_leaf_arteries == [1, 2]
    assert [a.param['R1'] for a in an.arteries[1:]] == [an.nondim['R1'][0]] * 2
    with pytest.raises(KeyError):
        af.ArteryNetwork(af.ParamParser('config/demo_arterybranch_geo.cfg'))
    p = af.ParamParser('config/demo_arterybranch_geo.cfg')
    p.param['R_term'] = 0.296
    an = af.ArteryNetwork(p)
    np.testing.assert_allclose(an.nondim['Ru'], [0.37, 0.296, 0.296])
    np.testing.assert_allclose(an.nondim['L'], [7.4, 5.92, 5.91])


def test_unit_conversions_like_the_reference_tests():
    """/root/reference/test/test_utils.py:8-17, same loops and constants."""
    from bloodflow_b200.arteryfe.utils import is_near, mmHg_to_unit, unit_to_mmHg
    for p in range(0, 30000, 500):
        assert is_near(unit_to_mmHg(p), p / 1333.22368421053)
    for p in range(0, 300, 5):
        assert is_near(mmHg_to_unit(p), p * 1333.223684210535)


def test_redimensionalise_inverts_nondimensionalise():
    """utils.py:107-170: the four natures and the pass-through; the values of the demo cfg
    (rc = 1, qc = 10, rho = 1.06) and of a non-unit scaling."""
    from bloodflow_b200.arteryfe.utils import nondimensionalise, redimensionalise
    x = np.array([0.0, 0.37, 4.1, 119990.131579])
    for rc, qc, rho in ((1.0, 10.0, 1.06), (1.1, 9.0, 1.06)):
        for nature, factor in (('time', qc / rc ** 3), ('area', 1 / rc ** 2), ('flow', 1 / qc),
                               ('pressure', rc ** 4 / rho / qc ** 2)):
            nd = nondimensionalise(rc, qc, rho, x, nature)
            np.testing.assert_allclose(nd, x * factor, rtol=1e-15)
            np.testing.assert_allclose(redimensionalise(rc, qc, rho, nd, nature), x, rtol=1e-15)
        np.testing.assert_array_equal(redimensionalise(rc, qc, rho, x, 'anything else'), x)
