"""Per-member cost of the order-8 tree ensemble members on one GPU (why N-GPU
weak scaling of perturbed members is not N x member 0)."""
import sys
import time

sys.path.insert(0, '.')
import bloodflow_b200.arteryfe as af                       # noqa: E402
from bloodflow_b200.ensemble import perturb_param          # noqa: E402
from bloodflow_b200.synthetic import tree_param            # noqa: E402

K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
for m in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    p = tree_param(8, Nx=1000, Nt=8000)
    if m:
        p = perturb_param(p, m)
    an = af.ArteryNetwork(p)
    an.define_x()
    an._dev.step(0, 100)
    an._dev.sync()
    t = time.perf_counter()
    an._dev.step(100, K)
    an._dev.sync()
    dt = time.perf_counter() - t
    tot = an._dev.get_totals()
    print('member %d: %.2f us/step, artery solves/step/artery %.3f, junction solves %.3f, picard %.1f' %
          (m, 1e6 * dt / K, tot[0] / (K + 100) / 255.0, tot[1] / (K + 100) / 127.0, tot[2] / (K + 100) / 128.0))
