"""Host-side logic of bench.py (no GPU): the launch count claimed per schedule and the member a rank runs."""
import importlib.util
import os
import types

import numpy as np

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location('bench_module', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_launch_count_per_schedule():
    b = _bench()
    Nt = 8000
    mask = (np.arange(Nt) % (Nt / 100) == 0).astype(np.uint8)          # the reference's dump rule, 100 per cycle
    # per-step schedules: two kernels per step, one dump kernel per snapshot (+ 3 flag / mailbox resets with the flags)
    assert b.launches_in(0, 5, 20, Nt, mask) == (40, 0, 20)
    assert b.launches_in(1, 100, 8000, Nt, mask) == (2 * 8000 + 100 + 3, 100, 8000)
    # persistent: chunks end at multiples of 64 steps and before a dump step
    n, dumps, chunks = b.launches_in(2, 5, 20, Nt, mask)
    assert (n, dumps, chunks) == (2 * 1 + 0 + 3, 0, 1)
    n, dumps, chunks = b.launches_in(2, 0, 160, Nt, mask)               # dumps at 0 and 80: [0,64) [64,80) [80,128) [128,160)
    assert (dumps, chunks, n) == (2, 4, 2 * 4 + 2 + 3)
    n, dumps, chunks = b.launches_in(2, 100, 8000, Nt, mask)
    assert dumps == 100 and n == 2 * chunks + dumps + 3 and 125 <= chunks <= 225


def test_ranks_run_replicas_unless_asked_otherwise():
    b = _bench()
    assert [b.member_of(types.SimpleNamespace(tree_members='replicas'), r) for r in range(4)] == [0, 0, 0, 0]
    assert [b.member_of(types.SimpleNamespace(tree_members='perturbed'), r) for r in range(4)] == [0, 1, 2, 3]
