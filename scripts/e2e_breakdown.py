"""Where the end-to-end time of ArteryNetwork(param).solve(n_steps=K) goes
(host wall clock per stage; run on a GPU box)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bloodflow_b200.arteryfe as af          # noqa: E402
from bloodflow_b200 import _ffi               # noqa: E402
from bloodflow_b200.synthetic import tree_param   # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for rep in range(3):
    t = [time.perf_counter()]
    param = tree_param(8, Nx=1000, Nt=8000)
    param.geo['N_cycles'] = 1
    param.solution['N_cycles_store'] = 1
    an = af.ArteryNetwork(param)
    t.append(time.perf_counter())
    an.define_x()
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
    an._dev.sync()
    t.append(time.perf_counter())
    an._dev.close()
    t.append(time.perf_counter())
    names = ('construct', 'set_store', 'stream_to_host(pinned ring)', 'step(enqueue)', 'fetch(drain)', 'sync', 'close')
    print('rep %d K=%d snapshots=%d: ' % (rep, K, n_snap) +
          ', '.join('%s %.1f ms' % (n, 1e3 * (b - a)) for n, a, b in zip(names, t[:-1], t[1:])) +
          ' | total %.1f ms' % (1e3 * (t[-1] - t[0])))
