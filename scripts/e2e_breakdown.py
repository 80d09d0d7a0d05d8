"""Where the end-to-end time of ArteryNetwork(param).solve(n_steps=K) goes: host wall clock per stage of
the same calls solve() makes, several repetitions (run on a GPU box).

    python scripts/e2e_breakdown.py [K] [repetitions]"""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bloodflow_b200.arteryfe as af          # noqa: E402
from bloodflow_b200 import _ffi               # noqa: E402
from bloodflow_b200.synthetic import tree_param   # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for rep in range(REPS):
    param = tree_param(8, Nx=1000, Nt=8000)
    param.geo['N_cycles'] = 1
    param.solution['N_cycles_store'] = 1
    t = [time.perf_counter()]
    an = af.ArteryNetwork(param)
    t.append(time.perf_counter())
    an.define_x()
    an._dev.clear_flags()
    t.append(time.perf_counter())
    Nt = an.geo['Nt']
    mask = an.store_mask()
    g = np.arange(K)
    n_snap = int(np.count_nonzero(mask[g % Nt] != 0))
    fields = _ffi.AF_STORE_FLOW | _ffi.AF_STORE_AREA | _ffi.AF_STORE_PRESSURE
    an._dev.set_store(mask, 0, max(1, n_snap), fields)
    t.append(time.perf_counter())
    an._dev.stream_store_to_host()
    t.append(time.perf_counter())
    an._dev.step(0, K)
    t.append(time.perf_counter())
    snaps = an._dev.fetch_snapshots()
    t.append(time.perf_counter())
    bad = an._dev.failed_step()
    flags = an._dev.get_flags()
    t.append(time.perf_counter())
    t_total = t[-1] - t[0]
    del snaps
    an._dev.close()
    t.append(time.perf_counter())
    names = ('construct', 'define_x+clear_flags', 'set_store', 'host result arrays', 'step(enqueue)', 'fetch(drain)',
             'failed_step+flags', 'close (not in e2e)')
    print('rep %d K=%d snapshots=%d: ' % (rep, K, n_snap) +
          ', '.join('%s %.2f' % (n, 1e3 * (b - a)) for n, a, b in zip(names, t[:-1], t[1:])) +
          ' | e2e %.2f ms' % (1e3 * t_total), flush=True)
# and the public call itself
for rep in range(REPS):
    param = tree_param(8, Nx=1000, Nt=8000)
    param.geo['N_cycles'] = 1
    param.solution['N_cycles_store'] = 1
    t0 = time.perf_counter()
    an = af.ArteryNetwork(param)
    t1 = time.perf_counter()
    res = an.solve(write_output=False, n_steps=K)
    t2 = time.perf_counter()
    print('solve() rep %d: construct %.2f ms, solve %.2f ms' % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
    del res, an
