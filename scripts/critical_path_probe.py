"""Which part of a time step of the order-8 tree is on the critical path: the same K steps with the
boundary work cut down (results are then wrong -- timing only).

    python scripts/critical_path_probe.py [K]

Variants: junction Newton limited to 1 / 2 solves, Windkessel Picard limited to 1 iteration, and the pure
chain with both.  A variant that does not shorten the step is not on the critical path."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bloodflow_b200.arteryfe as af                      # noqa: E402
from bloodflow_b200.engine import DeviceNetwork           # noqa: E402
from bloodflow_b200.synthetic import tree_param           # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
an = af.ArteryNetwork(tree_param(8, Nx=1000, Nt=8000, theta=0.6))
d = an._description


def device(**constants):
    return DeviceNetwork(
        d['Nx'], d['Nt'], d['theta'], d['dt'], d['junctions'], d['leaves'],
        k1=d['k1'], k2=d['k2'], k3=d['k3'], rho=d['rho'], Re=d['Re'], db=d['db'], p0=d['p0'],
        Ru=d['Ru'][None], Rd=d['Rd'][None], L=d['L'][None], q0=d['q0'][None],
        R1=d['R1'][None], R2=d['R2'][None], CT=d['CT'][None], q_ins=d['q_ins'], **constants)


an.define_x()
x0 = an._dev.get_junction_x()
for name, constants in (('as shipped', {}), ('junction_k_max=2', dict(junction_k_max=2)),
                        ('junction_k_max=1', dict(junction_k_max=1)), ('picard_k_max=1', dict(picard_k_max=1)),
                        ('junction_k_max=1, picard_k_max=1', dict(junction_k_max=1, picard_k_max=1)),
                        ('newton_rtol=1 (artery: one solve, test passes at once)', dict(newton_rtol=1.0e30))):
    dev = device(**constants)
    dev.set_junction_x(x0)
    dev.step(0, 200)
    dev.sync()
    t0 = time.perf_counter()
    dev.step(200, K)
    dev.sync()
    dt = time.perf_counter() - t0
    tot = dev.get_totals()
    print('%-60s %7.2f us/step   artery solves/step %.2f  junction solves/step %.2f  picard its/call %.1f  flags 0x%x'
          % (name, dt / K * 1e6, tot[0] / (K + 200.0) / dev.na, tot[1] / (K + 200.0) / max(dev.nj, 1),
             tot[2] / (K + 200.0) / max(dev.nl, 1), int(dev.get_flags()[0])), flush=True)
    dev.close()
