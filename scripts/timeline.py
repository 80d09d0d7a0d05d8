"""Per-task %globaltimer timeline of the order-8 tree's time loop (debug build of the library).

    make -C bloodflow_b200/csrc timeline
    ARTERYFE_B200_LIB=bloodflow_b200/csrc/libarteryfe_b200_tl.so AF_TIMELINE_STEPS=64 python scripts/timeline.py [K] [theta] [member]

Stamps (ns): artery 0 start, 1 own flag seen, 2 first-pass cells done, 3 Dirichlet flags seen, 4 first norm,
5 last solve + update done, 6 converged, 7 released; junction / leaf 0 start, 1 flags seen, 2 set-up done,
3 Newton / Picard done, 4 released.  Prints mean phase durations over the last 64 steps, the mean hand-over
latencies (release -> seen) and where the slowest chain of a step spends its time."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, '.')
import bloodflow_b200.arteryfe as af                      # noqa: E402
from bloodflow_b200 import _ffi                            # noqa: E402
from bloodflow_b200.synthetic import tree_param           # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
theta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
NS = int(os.environ.get('AF_TIMELINE_STEPS', '64'))
member = int(sys.argv[3]) if len(sys.argv) > 3 else 0
_p = tree_param(8, Nx=1000, Nt=8000, theta=theta)
if member:
    from bloodflow_b200.ensemble import perturb_param
    _p = perturb_param(_p, member)
an = af.ArteryNetwork(_p)
an.define_x()
dev = an._dev
K = (K // NS) * NS
dev.step(0, K)
dev.sync()
na, nj, nl = dev.na, dev.nj, dev.nl
nt = na + nj + nl
buf = np.zeros(NS * nt * 8, dtype=np.uint64)
fn = _ffi.lib.af_net_debug_timeline
fn.argtypes = [_ffi.Handle, ctypes.c_void_p, ctypes.c_int64]
_ffi.check(fn(dev._h, buf.ctypes.data, buf.size))
t = buf.reshape(NS, nt, 8).astype(np.int64)          # step g sits in row g % NS; K is a multiple of NS => row = order
art, jun, leaf = t[:, :na], t[:, na:na + nj], t[:, na + nj:]
if os.path.isdir('gpurun_out'):
    np.save('gpurun_out/timeline_raw.npy', t)
d = an._description
jn = np.asarray(d['junctions'])                        # [nj, 3] artery indices p, d1, d2
lv = np.asarray(d['leaves'])


def us(x):
    return float(np.mean(x)) / 1e3


step_ns = np.diff(art[:, :, 7].max(axis=1))
print('steps timed %d, mean step %.2f us (last release to last release)' % (NS, us(step_ns)))
print('artery   (mean over %d arteries x %d steps, us):' % (na, NS))
names = ('start->own flag', 'own flag->cells done', 'cells done->Dirichlet flags seen', 'flags seen->first norm',
         'first norm->last update', 'last update->converged', 'converged->released')
for k, nm in enumerate(names):
    print('   %-36s %6.2f' % (nm, us(art[:, :, k + 1] - art[:, :, k])))
print('   %-36s %6.2f' % ('life', us(art[:, :, 7] - art[:, :, 0])))
print('   %-36s %6.2f' % ('Dirichlet flags seen -> released', us(art[:, :, 7] - art[:, :, 3])))
print('junction (mean over %d x %d):' % (nj, NS))
for k, nm in enumerate(('start->flags seen', 'flags seen->set-up done', 'set-up->Newton done', 'Newton done->released')):
    print('   %-36s %6.2f' % (nm, us(jun[:, :, k + 1] - jun[:, :, k])))
print('   %-36s %6.2f' % ('flags seen -> released', us(jun[:, :, 4] - jun[:, :, 1])))
print('leaf (mean over %d x %d):' % (nl, NS))
for k, nm in enumerate(('start->flag seen', 'flag seen->stencil done', 'Picard', 'Picard done->released')):
    print('   %-36s %6.2f' % (nm, us(leaf[:, :, k + 1] - leaf[:, :, k])))
# hand-overs: junction j of step g sees its flags after the LAST of its three arteries released step g-1
rel_prev = art[:-1, :, 7]                               # release of step g-1
jseen = jun[1:, :, 1]
last3 = np.max(rel_prev[:, jn], axis=2)                 # [NS-1, nj]
print('hand-over artery release -> junction sees it : %.2f us (junction resident before: %.0f %%)'
      % (us(jseen - last3), 100.0 * np.mean(jun[1:, :, 0] < last3)))
# artery a of step g sees its Dirichlet flags after its junction(s)/leaf released
art_in = np.full(na, -1)
art_out = np.full(na, -1)
for j, (p, d1, d2) in enumerate(jn):
    art_out[p] = j
    art_in[d1] = j
    art_in[d2] = j
leaf_of = {int(a): k for k, a in enumerate(lv)}
rel_b = np.zeros((NS, na), dtype=np.int64)
for a in range(na):
    cands = []
    if art_in[a] >= 0:
        cands.append(jun[:, art_in[a], 4])
    if art_out[a] >= 0:
        cands.append(jun[:, art_out[a], 4])
    if a in leaf_of:
        cands.append(leaf[:, leaf_of[a], 4])
    rel_b[:, a] = np.max(np.stack(cands), axis=0)
print('hand-over boundary release -> artery sees it : %.2f us (artery had its cells done before: %.0f %%)'
      % (us(art[:, :, 3] - rel_b), 100.0 * np.mean(art[:, :, 2] < rel_b)))
print('artery start relative to the previous release of the SAME artery: %.2f us' % us(art[1:, :, 0] - art[:-1, :, 7]))
print('junction start relative to last release of its arteries (negative = resident early): %.2f us'
      % us(jun[1:, :, 0] - last3))
# slowest artery per step: which generation, how long after its flags
gen = np.floor(np.log2(np.arange(na) + 1)).astype(int)
worst = art[:, :, 7].argmax(axis=1)
print('last artery of a step by generation:', np.bincount(gen[worst], minlength=8).tolist())
print('per generation: mean (Dirichlet seen -> released) us:',
      [round(us((art[:, :, 7] - art[:, :, 3])[:, gen == g]), 2) for g in range(8)])
print('per generation: mean release time relative to the step\'s last release (us):',
      [round(us((art[:, :, 7] - art[:, :, 7].max(axis=1, keepdims=True))[:, gen == g]), 2) for g in range(8)])

# the critical cycle: root artery <-> root junction (mean over steps, us, relative to the root's previous release)
r_prev = art[:-1, 0, 7]
print('root artery, relative to its previous release:')
for k, nm in enumerate(('start', 'own flag seen', 'cells done', 'Dirichlet seen', 'first norm', 'last update', 'converged', 'released')):
    print('   %-16s %7.2f' % (nm, us(art[1:, 0, k] - r_prev)))
print('root junction, relative to the root artery\'s previous release:')
for k, nm in enumerate(('start', 'flags seen', 'set-up done', 'Newton done', 'mailbox written')):
    print('   %-16s %7.2f' % (nm, us(jun[1:, 0, k] - r_prev)))
for a_ in (1, 2):
    print('daughter %d released relative to root release (same step): %.2f us' % (a_, us(art[:, a_, 7] - art[:, 0, 7])))
for g in range(8):
    sel = gen == g
    print('gen %d: start - previous release of the same artery %.2f us; Dirichlet seen - cells done %.2f us; '
          'resident before own previous release: %.0f %%'
          % (g, us(art[1:, sel, 0] - art[:-1, sel, 7]), us(art[:, sel, 3] - art[:, sel, 2]),
             100.0 * np.mean(art[1:, sel, 0] < art[:-1, sel, 7])))

# junction Newton: stamps at the top of iterations 1 and 2 (slots 5, 6) and the iteration count at exit (slot 7)
k_exit = jun[:, :, 7]
print('junction Newton exits after k linearisations: histogram', np.bincount(k_exit.flatten().clip(0, 31))[:8].tolist())
ok = (jun[:, :, 5] > 0) & (jun[:, :, 6] > 0)
print('junction: set-up done -> top of iteration 1: %.2f us; iteration 1 -> 2: %.2f us'
      % (us((jun[:, :, 5] - jun[:, :, 2])[jun[:, :, 5] > 0]), us((jun[:, :, 6] - jun[:, :, 5])[ok])))
