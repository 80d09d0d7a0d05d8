"""Measured deviation of the CUDA path from (a) the reference-generated golden runs
(tests/golden/run_*.npz) and (b) the C oracle on the full-size synthetic trees C3 / C4 over a whole
cardiac cycle, every `--every` steps: max relative deviation of A, q and the junction unknowns, and
the number of artery / junction Newton counts that differ in that window.  Run on a GPU box; the
lines are committed under profiles/ and quoted in DESIGN.md section 2.

    python scripts/parity_report.py                      # golden cases c1, c2, o3
    python scripts/parity_report.py --config c4 --theta 0.6 [--steps 8000] [--every 500]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bloodflow_b200.arteryfe as af      # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
SHAPES = {'c3': (5, 500, 4000), 'c4': (8, 1000, 8000)}


def golden():
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
        print(json.dumps(out), flush=True)


def tree(config, theta, steps, every):
    from bloodflow_b200.synthetic import tree_param
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle
    order, Nx, Nt = SHAPES[config]
    steps = Nt if steps is None else steps
    an = af.ArteryNetwork(tree_param(order, Nx=Nx, Nt=Nt, theta=theta))
    an.define_x()
    dev = an._dev
    dev.set_counter_log(steps)
    p = tree_param(order, Nx=Nx, Nt=Nt, theta=theta)
    c = COracle(ao.OracleNetwork(p.param, p.geo, p.solution))
    done, tot_a, tot_j = 0, 0, 0
    while done < steps:
        n = min(every, steps - done)
        dev.step(done, n)
        failed, la, lj = c.step(done, n, nthreads=0, log=True)
        A, q = dev.get_state()
        Ao, qo = c.state()
        art, jun, _ = dev.fetch_counter_log()
        ma = int(np.count_nonzero(art[done:done + n, 0] != la))
        mj = int(np.count_nonzero(jun[done:done + n, 0] != lj))
        tot_a += ma
        tot_j += mj
        done += n
        qs = np.abs(qo).max()
        print(json.dumps({
            "config": config, "theta": theta, "step": done,
            "max_rel_A": float(np.abs(A[0] / Ao - 1).max()),
            "max_rel_q": float((np.abs(q[0] - qo) / np.maximum(np.abs(qo), 1e-3 * qs)).max()),
            "max_rel_junction_x": float(np.abs(dev.get_junction_x()[0] / c.x() - 1).max()),
            "artery_count_mismatches": ma, "of_artery_solves": int(la.size),
            "junction_count_mismatches": mj, "of_junction_solves": int(lj.size),
            "oracle_failed": int(failed), "flags": int(dev.get_flags()[0])}), flush=True)
    print(json.dumps({"config": config, "theta": theta, "steps": steps, "artery_count_mismatches_total": tot_a,
                      "junction_count_mismatches_total": tot_j}), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='golden', choices=('golden', 'c3', 'c4'))
    ap.add_argument('--theta', type=float, default=0.6)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--every', type=int, default=500)
    a = ap.parse_args()
    if a.config == 'golden':
        golden()
    else:
        tree(a.config, a.theta, a.steps, a.every)
