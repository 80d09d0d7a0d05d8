"""Perturbation growth of the reference scheme on a synthetic tree (C oracle against a copy of itself
whose state differs by a relative 1e-13 in one vertex): max relative difference of A every `every` steps.
Used to choose the C3/C4 bench variants ("confirm stability at orders 5/8 and adjust").

    python scripts/stability_probe.py --order 5 --Nx 500 --Nt 4000 --theta 0.55 --steps 4000
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bloodflow_b200.synthetic import tree_param      # noqa: E402
from oracle import arteryfe_oracle as ao             # noqa: E402
from oracle.oracle_c import COracle                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--order', type=int, default=5)
ap.add_argument('--Nx', type=int, default=500)
ap.add_argument('--Nt', type=int, default=4000)
ap.add_argument('--theta', type=float, default=0.55)
ap.add_argument('--L', type=float, default=20.0)
ap.add_argument('--alpha', type=float, default=0.8)
ap.add_argument('--steps', type=int, default=4000)
ap.add_argument('--every', type=int, default=200)
ap.add_argument('--eps', type=float, default=1e-13)
args = ap.parse_args()


def make():
    p = tree_param(args.order, Nx=args.Nx, Nt=args.Nt, theta=args.theta, L=args.L, alpha=args.alpha)
    return COracle(ao.OracleNetwork(p.param, p.geo, p.solution))


a, b = make(), make()
A, q = b.state()
A[min(3, a.na - 1), min(100, a.np - 1)] *= 1 + args.eps
b.set_state(A, q)
done = 0
print(json.dumps({"config": vars(args)}))
while done < args.steps:
    n = min(args.every, args.steps - done)
    fa = a.step(done, n, nthreads=0)
    fb = b.step(done, n, nthreads=0)
    done += n
    Aa, qa = a.state()
    Ab, qb = b.state()
    pa = a.pressure()
    print(json.dumps({"step": done, "max_rel_A": float(np.abs(Ab / Aa - 1).max()),
                      "max_abs_q": float(np.abs(qb - qa).max()), "q_scale": float(np.abs(qa).max()),
                      "p_mmHg_min": float(pa.min() / 1333.22365 * 1060 * 100 / 1.0 ** 4 if False else pa.min()),
                      "p_max": float(pa.max()), "failed": int(fa or fb),
                      "q_roughness": float(np.abs(np.diff(qa, 2, axis=1)).max())}), flush=True)
    if fa or fb or not np.isfinite(Aa).all():
        break
