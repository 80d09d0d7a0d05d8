"""Measured deviation of the CUDA path from the reference-generated golden runs
(tests/golden/run_*.npz): max relative error of every stored field over the whole
run, and the Newton-count agreement.  Run on a GPU box; the numbers are quoted in
DESIGN.md section 2."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bloodflow_b200.arteryfe as af      # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for case in ('c1', 'c2', 'o3'):
    g = np.load(os.path.join(GOLDEN, 'run_%s.npz' % case))
    an = af.ArteryNetwork(af.ParamParser(os.path.join(GOLDEN, 'golden_%s.cfg' % case)))
    n_steps = an.geo['Nt'] * an.geo['N_cycles']
    an._dev.set_counter_log(n_steps)
    res = an.solve(write_output=False)
    out = {"case": case, "steps": int(n_steps), "snapshots": int(len(res['t']))}
    for field in ('flow', 'area', 'pressure'):
        scale = np.abs(g[field]).max()
        out["max_rel_" + field] = float(np.max(np.abs(res[field] - g[field]) / np.maximum(np.abs(g[field]), 1e-3 * scale)))
    out["max_rel_junction_x"] = float(np.max(np.abs(np.asarray(an.x) - g['x_final']) / np.abs(g['x_final'])))
    art_its, j_its, p_its = an._dev.fetch_counter_log()
    out["artery_newton_counts_equal"] = bool(np.array_equal(art_its[:, 0, :], g['artery_its']))
    if 'junction_its' in g.files:
        out["junction_newton_counts_equal"] = bool(np.array_equal(j_its[:, 0, :], g['junction_its']))
    print(json.dumps(out))
