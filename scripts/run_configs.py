#!/usr/bin/env python
"""Run the BASELINE.json configurations C1..C5 through the public API on one
GPU and print one JSON line per config (throughput, iteration statistics, flags,
pressure ranges, cycle-to-cycle periodicity).  Used to fill profiles/ and to
check the stability of the synthetic trees over full cardiac cycles."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

import bloodflow_b200.arteryfe as af                       # noqa: E402
from bloodflow_b200.arteryfe.utils import unit_to_mmHg, redimensionalise   # noqa: E402
from bloodflow_b200.synthetic import tree_param             # noqa: E402


def summarise(name, an, res, seconds, steps):
    dev = an._dev
    nodes = dev.na * dev.np
    tot = dev.get_totals()
    nd = an.nondim
    p = unit_to_mmHg(redimensionalise(nd['rc'], nd['qc'], nd['rho'], res['pressure'], 'pressure'))
    out = {
        "config": name, "arteries": dev.na, "vertices": nodes, "steps": steps, "seconds": seconds,
        "node_steps_per_s": nodes * steps / seconds, "flags": int(dev.get_flags()[0]),
        "artery_newton_its_per_solve": tot[0] / float(dev.na * steps),
        "junction_solves_per_step": tot[1] / float(max(dev.nj, 1) * steps),
        "picard_its_per_call": tot[2] / float(max(dev.nl, 1) * steps),
        "snapshots": int(len(res['t'])), "p_mmHg_min": float(np.nanmin(p)), "p_mmHg_max": float(np.nanmax(p)),
        "q_min": float(np.nanmin(res['flow'])), "q_max": float(np.nanmax(res['flow'])),
        "finite": bool(np.isfinite(res['flow']).all()),
    }
    return out


def run(name, param, **kw):
    an = af.ArteryNetwork(param, **kw)
    steps = an.geo['Nt'] * an.geo['N_cycles']
    t0 = time.perf_counter()
    try:
        res = an.solve(write_output=False)
    except RuntimeError as e:
        print(json.dumps({"config": name, "error": str(e)}), flush=True)
        return None
    dt = time.perf_counter() - t0
    out = summarise(name, an, res, dt, steps)
    print(json.dumps(out), flush=True)
    return an, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c1,c1head,c2,c3,c4,c5')
    ap.add_argument('--members', type=int, default=4096)
    args = ap.parse_args()
    want = args.configs.split(',')
    if 'c1' in want:
        p = af.ParamParser('config/demo_arterybranch.cfg')
        run('C1 demo_arterybranch.cfg, wk_scale=(1,1), 4 cycles', p, wk_scale=(1, 1))
    if 'c1head' in want:
        p = af.ParamParser('config/demo_arterybranch.cfg')
        run('C1 demo_arterybranch.cfg, HEAD factors (0.3,0.01)', p)
    if 'c2' in want:
        p = af.ParamParser('config/demo_arterybranch_geo.cfg')
        p.param['R_term'] = 0.296
        run('C2 demo_arterybranch_geo.cfg + R_term=0.296, HEAD factors', p)
    if 'c3' in want:
        run('C3 order-5 tree Nx=500 Nt=4000, 2 cycles', tree_param(5, Nx=500, Nt=4000, N_cycles=2, Nt_store=100))
    if 'c4' in want:
        run('C4 order-8 tree Nx=1000 Nt=8000, 4 cycles', tree_param(8, Nx=1000, Nt=8000, N_cycles=4, Nt_store=100))
    if 'c5' in want:
        from bloodflow_b200 import ensemble as en
        base = tree_param(4, Nx=100, Nt=2000, N_cycles=2, Nt_store=100)
        t0 = time.perf_counter()
        ens = en.EnsembleNetwork(base, args.members)
        op = ens.solve(n_cycles=2)
        dt = time.perf_counter() - t0
        flags = ens.flags()
        nd = ens.nominal.nondim
        p = unit_to_mmHg(redimensionalise(nd['rc'], nd['qc'], nd['rho'], op, 'pressure'))
        tot = ens.dev.get_totals()
        steps = 2 * 2000
        print(json.dumps({
            "config": "C5 UQ ensemble, %d perturbed order-4 networks (Nx=100, Nt=2000), 2 cycles" % args.members,
            "members": args.members, "vertices": ens.vertices, "steps": steps, "seconds": dt,
            "node_steps_per_s": ens.vertices * steps / dt, "cycles_per_s_all_networks": args.members * 2 / dt,
            "members_with_artery_nonconvergence": int(np.count_nonzero(flags & 1)),
            "members_with_nonfinite": int(np.count_nonzero(flags & 8)),
            "artery_newton_its_per_solve": tot[0] / float(ens.n_networks * ens.dev.na * steps),
            "junction_solves_per_step": tot[1] / float(ens.n_networks * ens.dev.nj * steps),
            "outlet_p_mmHg_mean": float(p.mean()), "outlet_p_mmHg_std_over_members": float(p.std(axis=1).mean()),
            "outlet_p_mmHg_min": float(p.min()), "outlet_p_mmHg_max": float(p.max()), "snapshots": int(op.shape[0]),
        }), flush=True)


if __name__ == '__main__':
    main()
