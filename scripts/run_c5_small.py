"""A short run of the C5-shaped ensemble (order 4, Nx = 100, Nt = 2000) for ncu captures:
    python scripts/run_c5_small.py [members] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
from bloodflow_b200 import ensemble as en                # noqa: E402
from bloodflow_b200.synthetic import tree_param          # noqa: E402

members = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
ens = en.EnsembleNetwork(tree_param(4, Nx=100, Nt=2000, N_cycles=2, Nt_store=100), members)
ens.dev.step(0, steps)
ens.dev.sync()
print('members', members, 'steps', steps, 'vertices', ens.vertices, 'flagged', ens.flagged_members(),
      'totals', ens.dev.get_totals())
