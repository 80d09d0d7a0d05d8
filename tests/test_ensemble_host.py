"""Host-side logic of the UQ ensemble (no GPU): the vectorised member
descriptions equal building every member through the facade, the shard
partition covers the index range, and the moment reduction over two ranks
(gloo, CPU tensors -- the same code path NCCL runs on the GPUs) equals the
single-process statistics."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT


def test_stacked_description_equals_per_member_setup():
    from bloodflow_b200 import ensemble as en
    from bloodflow_b200.arteryfe.artery_network import ArteryNetwork
    from bloodflow_b200.synthetic import tree_param
    base = tree_param(4, Nx=100, Nt=2000, N_cycles=2)
    idx = [0, 3, 17, 65535]
    desc, nominal = en.stacked_description(base, idx)
    assert desc['Ru'].shape == (4, 15) and desc['R1'].shape == (4, 8)
    for m, i in enumerate(idx):
        one = ArteryNetwork(en.perturb_param(base, i), _host_only=True)._description
        for key in ('Ru', 'Rd', 'L', 'q0', 'R1', 'R2', 'CT'):
            np.testing.assert_allclose(desc[key][m], one[key], rtol=1e-14, err_msg=key)
        for key in ('k1', 'k3'):
            np.testing.assert_allclose(desc[key][m], one[key], rtol=1e-15)
        for key in ('k2', 'rho', 'Re', 'db', 'p0', 'dt', 'theta'):
            assert desc[key] == one[key]
        np.testing.assert_array_equal(desc['junctions'], one['junctions'])
        np.testing.assert_array_equal(desc['leaves'], one['leaves'])
    # worst-case Picard gain of the recipe stays below one
    assert desc['R1'].min() > 0


def test_shard_partition():
    from bloodflow_b200.ensemble import shard
    for n, w in ((65536, 8), (10, 3), (7, 8)):
        spans = [shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from bloodflow_b200 import ensemble as en
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo')
    rng = np.random.default_rng(7)
    op_all = torch.from_numpy(1100.0 + 30.0 * rng.standard_normal((5, 11, 8)))   # [snapshot, member, leaf]
    lo, hi = en.shard(op_all.shape[1], rank, world)
    stats = en.finalize(en.reduce_moments(en.local_moments(op_all[:, lo:hi].contiguous())))
    ref = en.finalize(en.local_moments(op_all))
    for k in ref:
        assert torch.allclose(stats[k], ref[k], rtol=1e-12, atol=1e-9), k
    assert float(stats['count'][0, 0]) == 11.0
    dist.barrier()
    dist.destroy_process_group()
    print('rank', rank, 'ok')
''')


def test_moment_reduction_two_ranks_gloo(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', '29533', str(script)],
        capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count('ok') == 2
