#!/usr/bin/env python
"""bench.py -- artery-node-timesteps/s of the per-timestep network solve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json config 4, the one the 1e9 target is quoted on): the
synthetic order-8 tree, 255 arteries x 1001 vertices, Nt = 8000 steps per cardiac cycle,
Windkessel outlets on the 128 leaves, data/example_inlet.csv.  A "step" is one time step
of the whole network (set_bcs + every artery solve).  N GPUs: N replicas of that tree, one
per GPU (weak scaling; a single network does not shard; --tree-members perturbed runs N
parameter-perturbed members instead, whose costs differ).

The same JSON line carries `ensemble_c5`: BASELINE.json config 5, 65 536 perturbed order-4
networks SHARDED over the N GPUs (strong scaling), exactly its two cardiac cycles, the
outlet-pressure statistics reduced on the device and exchanged with one NCCL all-gather.

One JSON line on stdout (rank 0).  `value` is timed with CUDA events on the stream the
kernels run on, inputs resident in HBM; `e2e` is the same metric through the public
arteryfe API (ArteryNetwork(param).solve()) with host buffers: upload of the network
description, the loop with its snapshots, download of the snapshots.

--impl reference: the reference CPU path.  FEniCS cannot run here, so this is the C
restatement oracle/oracle.c ("restated reference"), OpenMP over all host cores, same
config and metric; nothing of the CUDA library is loaded on that path.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.chdir(ROOT)

METRIC = "artery-node-timesteps/sec"
UNIT = "node-steps/s"
ORDER, NX, NT = 8, 1000, 8000
THETA = 0.6          # synthetic trees: the variant on which the scheme does not amplify round-off
NT_STORE = 100
C5_MEMBERS, C5_ORDER, C5_NX, C5_NT, C5_CYCLES = 65536, 4, 100, 2000, 2
# algorithmic FP64 work per vertex-step: residual 110, Jacobian 140, block-Thomas solve 50
W_RES, W_JAC, W_SOLVE = 110.0, 140.0, 50.0


def flops_per_vertex_step(k):
    return (k + 1.0) * W_RES + k * (W_JAC + W_SOLVE)


def tree_member(order, Nx, Nt, member, N_cycles=4, theta=THETA):
    """Member `member` of the UQ ensemble around the nominal tree (SURVEY 8(d) perturbation model);
    member 0 is the nominal network.  Pure host code (no CUDA library)."""
    from bloodflow_b200.synthetic import perturb_param, tree_param
    p = tree_param(order, Nx=Nx, Nt=Nt, N_cycles=N_cycles, Nt_store=NT_STORE, theta=theta)
    return perturb_param(p, member) if member else p


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the run (NVML)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self._active = threading.Event()     # NVML queries contend with the CUDA driver: sample only when asked
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop_evt.is_set():
            if not self._active.is_set():
                self._active.wait(0.05)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def mark(self):
        return len(self.samples)

    def resume(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def summary(self, lo=0, hi=None):
        s = self.samples[lo:hi]
        return {"sm_mhz": float(np.median(s)) if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()


def committed_capture(kernel):
    """Numbers of the dominant kernel from the committed `ncu --set full` capture of this round
    (profiles/r02_traffic.json): DRAM bytes per launch and pipe utilisation.  None if absent."""
    for name in ('r02_traffic.json', 'r01_traffic.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as fh:
                t = json.load(fh)[kernel]
            t = dict(t)
            t['source'] = 'profiles/' + name
            return t
        except Exception:
            continue
    return None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------
# CPU side: the restated reference (oracle/oracle.c).  Only these two functions touch oracle/.
# ---------------------------------------------------------------------------------------
def _c_oracle(order, Nx, Nt, member, theta):
    from oracle import arteryfe_oracle as ao
    from oracle.oracle_c import COracle
    p = tree_member(order, Nx, Nt, member, theta=theta)
    return COracle(ao.OracleNetwork(p.param, p.geo, p.solution))


def cpu_baseline(order, Nx, Nt, theta, budget_s=12.0):
    """C oracle on all host cores over a bounded sample of the same workload."""
    c = _c_oracle(order, Nx, Nt, 0, theta)
    cores = os.cpu_count() or 1
    nodes = c.na * c.np
    t0 = time.perf_counter()
    c.step(0, 4, nthreads=cores)
    per = (time.perf_counter() - t0) / 4
    n = int(max(8, min(4000, budget_s / per)))
    t0 = time.perf_counter()
    failed = c.step(4, n, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": nodes * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d time steps of the same order-%d tree after 4 warm-up steps, oracle/oracle.c with "
                      "OpenMP over arteries/junctions (%s)" % (n, order, "ok" if not failed else "FAILED"),
            "ms_per_step": dt / n * 1e3}


def run_reference(args, rank, world):
    """--impl reference.  Rank 0 alone works.  With N > 1 the arm times the SAME N members the GPU
    arm runs (one after the other on the one host), so the ratio stays like for like."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W, K = max(args.warmup, 0), args.steps
    n_members = max(world, args.gpus, 1)
    budget = 150.0 / n_members
    total_nodes_steps, total_s, failed, K_run = 0.0, 0.0, 0, K
    for m in range(n_members):
        c = _c_oracle(args.order, args.Nx, args.Nt, member_of(args, m), args.theta)
        nodes = c.na * c.np
        t0 = time.perf_counter()
        c.step(0, max(W, 1), nthreads=cores)
        per = (time.perf_counter() - t0) / max(W, 1)
        K_run = K if per * K <= budget else max(10, int(budget / per))      # bounded sample
        t0 = time.perf_counter()
        failed |= c.step(max(W, 1), K_run, nthreads=cores)
        dt = time.perf_counter() - t0
        total_nodes_steps += nodes * K_run
        total_s += dt
        c.close()
    value = total_nodes_steps / total_s
    sample = ("%d of the requested %d time steps of each of the %d order-%d tree(s) the GPU arm runs, "
              "one after the other (bounded sample), oracle/oracle.c, OpenMP %d threads"
              % (K_run, K, n_members, args.order, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K_run, "warmup": W, "ms_per_step": total_s / (K_run * n_members) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference CPU path = C restatement of the reference scheme (FEniCS 2017.2.0 cannot be "
                "installed here); a host does not get faster with N, so the aggregate over N members equals the "
                "one-member rate; failed=%d" % failed,
    }
    print(json.dumps(line), flush=True)


def member_of(args, rank):
    """Which member of the tree ensemble GPU `rank` runs: the nominal tree everywhere (replicas: a single
    network does not shard, so N GPUs measure N independent copies of the same work), or -- `--tree-members
    perturbed` -- member `rank` of the UQ ensemble around it (members then differ in cost by up to 16 %,
    profiles/r02_final_member_cost_*.log, and the slowest sets the line)."""
    return rank if args.tree_members == 'perturbed' else 0


def workload_config(args):
    return {"workload": "synthetic symmetric tree order=%d (%d arteries x %d vertices), Nt=%d per cycle, "
                        "Windkessel outlets on %d leaves, data/example_inlet.csv; N GPUs = %s, one per GPU (a single "
                        "network does not shard; the sharded workload is ensemble_c5)"
                        % (args.order, 2 ** args.order - 1, args.Nx + 1, args.Nt, 2 ** (args.order - 1),
                           "N perturbed members of the tree ensemble" if args.tree_members == 'perturbed'
                           else "N replicas of this tree"),
            "order": args.order, "Nx": args.Nx, "Nt": args.Nt, "theta": args.theta,
            "l2": "no flush: the time loop is a dependent chain (step n+1 reads exactly what step n wrote), so "
                  "L2 residency of the %.1f MB state is the real operating point; nothing is cached or skipped "
                  "across timed steps" % ((2 ** args.order - 1) * (args.Nx + 1) * 32 / 1e6),
            "parallelism": "ensemble-dp%d" % args.gpus}


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def phase_shares(dev, stream, g0, n, Nt, torch):
    """Replay of steps [g0, g0+n) with an event between the two kernels of every step: their shares of a
    step.  (The events defeat the programmatic-dependent-launch overlap, so the absolute times are longer
    than in the timed loop; only the SHARE is used, scaled by the timed ms_per_step.)"""
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    for i in range(n):
        g = g0 + i
        evs[i][0].record(stream)
        dev.phase_boundary()
        evs[i][1].record(stream)
        dev.phase_arteries((g % Nt + 1) % Nt)
        evs[i][2].record(stream)
    dev.sync()
    b = np.array([e[0].elapsed_time(e[1]) for e in evs]) * 1e3
    a = np.array([e[1].elapsed_time(e[2]) for e in evs]) * 1e3
    return {"boundary_us_serialised": float(np.mean(b)), "artery_us_serialised": float(np.mean(a)),
            "artery_share": float(np.mean(a) / (np.mean(a) + np.mean(b))), "replayed_steps": n}


def launches_in(dev_schedule, W, K, Nt, mask, chunk=64):
    """Kernels the library enqueues for af_net_step(W, K) on a single network: the per-step schedules launch the
    boundary and the artery kernel every step; the persistent schedule one pair per chunk (chunks end at multiples
    of `chunk` steps and before a step whose Un is dumped); plus one dump kernel per snapshot and the three
    kernels that re-base the flags / empty the mailbox at the start of the call."""
    due = np.tile(mask, 2 + (W + K) // Nt)[W:W + K] != 0
    n_dumps = int(due.sum())
    if dev_schedule != 2:
        return 2 * K + n_dumps + (3 if dev_schedule == 1 else 0), n_dumps, K
    chunks, g = 0, W
    while g < W + K:
        n = 1
        while (g + n) % chunk != 0 and g + n < W + K and not due[g + n - W]:
            n += 1
        chunks += 1
        g += n
    return 2 * chunks + n_dumps + 3, n_dumps, chunks


def end_to_end(args, rank, local_rank, K):
    import bloodflow_b200.arteryfe as af
    n_cycles = -(-K // args.Nt)
    p = tree_member(args.order, args.Nx, args.Nt, member_of(args, rank), N_cycles=n_cycles, theta=args.theta)
    p.solution['N_cycles_store'] = n_cycles          # snapshots from the first step on
    t0 = time.perf_counter()
    an = af.ArteryNetwork(p, device=local_rank)
    res = an.solve(write_output=False, n_steps=K)
    dt = time.perf_counter() - t0
    d2h = sum(res[k].nbytes for k in ('flow', 'area', 'pressure') if k in res)
    h2d = an.q_ins.nbytes + an._dev.na * 4 * 8 + an._dev.nl * 3 * 8 + 7 * 8 + (an._dev.nj * 3 + an._dev.nl) * 4
    return {"seconds": dt, "h2d": h2d, "d2h": d2h, "flags": an.flags}


def fp64_peak(device):
    import ctypes
    from bloodflow_b200 import _ffi
    v = ctypes.c_double(0)
    _ffi.check(_ffi.lib.af_measure_fp64_peak(device, ctypes.byref(v)))
    return v.value


def run_c5(args, rank, local_rank, world, torch, dist, sampler):
    """BASELINE config 5 in full: 65 536 parameter-perturbed order-4 networks sharded over the N GPUs
    (strong scaling), two cardiac cycles, outlet-pressure moments by the library's kernels and ONE
    all-gather.  Returns the record (rank 0) or None."""
    from bloodflow_b200 import ensemble as en
    from bloodflow_b200.synthetic import tree_param
    total = args.c5_members
    lo, hi = en.shard(total, rank, world)
    M = hi - lo
    steps = args.c5_steps
    n_cycles = -(-steps // C5_NT)
    base = tree_param(C5_ORDER, Nx=C5_NX, Nt=C5_NT, N_cycles=n_cycles, Nt_store=100)
    stream = torch.cuda.Stream(device=local_rank)
    # warm-up launches (module load, clocks) on a small throwaway ensemble of the same kernels, so that the
    # timed run below is the whole of C5 from its step 0
    warm = en.EnsembleNetwork(base, min(M, 9472), first_index=lo, device=local_rank, stream=stream.cuda_stream)
    warm.configure_store(1, 1)
    warm.dev.step(0, 25)
    with torch.cuda.stream(stream):
        en.combine_moments(en.device_moments(warm.dev, local_rank))      # loads the statistics kernels too
    torch.cuda.synchronize()
    warm.dev.close()
    t_setup = time.perf_counter()
    ens = en.EnsembleNetwork(base, M, first_index=lo, device=local_rank, stream=stream.cuda_stream)
    t_setup = time.perf_counter() - t_setup
    dev = ens.dev
    ens.configure_store(n_cycles, 1)
    nodes_total = total * dev.na * dev.np
    ctas, fill = M * dev.na, 2 * 148 * 16                       # member shares chosen by the library (AF_SUBSTREAMS)
    n_sub = 4 if ctas >= 4 * fill else (2 if ctas >= 2 * fill else 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    mark = sampler.mark()
    sampler.resume()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    dev.step(0, steps)
    e1.record(stream)
    barrier()
    sampler.pause()
    clocks = sampler.summary(mark)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # statistics: library kernels + one collective, timed on the device
    with torch.cuda.stream(stream):
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record(stream)
        packed = en.device_moments(dev, local_rank)
        c1.record(stream)
        stats = en.combine_moments(packed)
        c2.record(stream)
    torch.cuda.synchronize()
    flags = dev.get_flags()
    # one SUM all-reduce: members with a failure, the three iteration totals, and per status bit the number of
    # members that have it set (NCCL has no bitwise OR)
    bits = [int(np.count_nonzero(flags & np.uint32(1 << k))) for k in range(8)]
    packed_i = torch.tensor([int(np.count_nonzero(flags & np.uint32(1 | 8)))] + [int(v) for v in dev.get_totals()] + bits,
                            dtype=torch.int64, device='cuda')
    if world > 1:
        dist.all_reduce(packed_i, op=dist.ReduceOp.SUM)
    packed_i = packed_i.cpu().numpy()
    n_flagged, tot, bits = int(packed_i[0]), packed_i[1:4], packed_i[4:]
    flags_or = int(sum(1 << k for k in range(8) if bits[k] > 0))
    if rank != 0:
        return None
    peaks, peak_kind = measured_peaks()
    value = nodes_total * steps / (ms_total * 1e-3)
    its = float(tot[0]) / (float(total) * dev.na * steps)
    fl = flops_per_vertex_step(its)
    cap = committed_capture("k_artery<32>")
    peak_tf = fp64_peak(local_rank)
    return {
        "workload": "UQ ensemble (BASELINE config 5): %d perturbed order-%d networks (15 arteries x %d vertices, "
                    "Nt=%d), sharded over %d GPU(s), %d steps = %.2f cardiac cycles"
                    % (total, C5_ORDER, C5_NX + 1, C5_NT, world, steps, steps / C5_NT),
        "members_total": total, "members_per_gpu": M, "n_gpus": world, "steps": steps, "scaling": "strong",
        "value": value, "unit": UNIT, "ms_per_step": ms_total / steps, "seconds": ms_total * 1e-3,
        "cycles_per_sec_all_networks": total * (steps / C5_NT) / (ms_total * 1e-3),
        "flags_or": flags_or, "flagged_members": n_flagged,
        "members_per_status_bit": {"artery_noconv": int(bits[0]), "junction_kmax": int(bits[1]), "picard_kmax": int(bits[2]),
                                   "nonfinite": int(bits[3]), "singular": int(bits[4]), "picard_cycle": int(bits[5]),
                                   "sync_timeout": int(bits[6])},
        "flagged_members_meaning": "members with an artery Newton failure or a non-finite residual (must be 0)",
        "newton_its_per_artery_step": its,
        "collective": {"kind": "one all_gather of the packed moments [5, snapshots, leaves] per rank + fold kernel",
                       "moments_kernel_ms": c0.elapsed_time(c1), "all_gather_and_fold_ms": c1.elapsed_time(c2),
                       "bytes_per_rank": int(packed.numel() * 8)},
        "statistics": {"mean_outlet_pressure": float(stats['mean'].mean().item()),
                       "std_outlet_pressure": float(stats['var'].mean().sqrt().item()),
                       "min": float(stats['min'].min().item()), "max": float(stats['max'].max().item()),
                       "members_counted": float(stats['count'].flatten()[0].item())},
        "roofline": {"bound": "fp64", "achieved": fl * value / world / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": fl * value / world / 1e12 / peak_tf, "flops_per_vertex_step": fl, "per": "GPU",
                     "hbm": {"achieved": 32.0 * value / world / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": 32.0 * value / world / 1e9 / peaks["hbm_gbs"], "peak_kind": peak_kind,
                             "algorithmic_bytes_per_vertex_step": 32},
                     "ncu": cap},
        "gpu_launches": 2 * steps * n_sub + 100 + 2, "member_shares_per_step": n_sub, "clocks": clocks,
        "setup_seconds": t_setup, "state_mb_per_gpu": M * dev.na * dev.np * 32 / 1e6,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8000, help='timed time steps (default: one cardiac cycle)')
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--order', type=int, default=ORDER)
    ap.add_argument('--Nx', type=int, default=NX)
    ap.add_argument('--Nt', type=int, default=NT)
    ap.add_argument('--theta', type=float, default=THETA)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-c5', action='store_true', help='skip the ensemble_c5 record')
    ap.add_argument('--tree-members', choices=('replicas', 'perturbed'), default='replicas',
                    help='N > 1: every GPU runs the nominal tree (default) or member `rank` of the perturbed ensemble')
    ap.add_argument('--c5-members', type=int, default=C5_MEMBERS)
    ap.add_argument('--c5-steps', type=int, default=C5_CYCLES * C5_NT)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the "
                         "CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    import bloodflow_b200.arteryfe as af
    from bloodflow_b200 import _ffi

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.resume()                        # from before the warm-up on: short timed regions still get samples
    p = tree_member(args.order, args.Nx, args.Nt, member_of(args, rank), theta=args.theta)
    # the library enqueues on a torch stream, so torch CUDA events time exactly its kernels
    stream = torch.cuda.Stream(device=local_rank)
    an = af.ArteryNetwork(p, device=local_rank, stream=stream.cuda_stream)
    dev = an._dev
    schedule = dev.schedule()
    nodes = dev.na * dev.np
    an.define_x()
    mask = an.store_mask()
    dev.set_store(mask, 0, max(1, (K + W) // (args.Nt // NT_STORE) + 2), _ffi.AF_STORE_OUTLET_PRESSURE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed loop: W warm-up steps, then exactly K steps ----------
    dev.step(0, W)
    barrier()
    saved = (dev.get_state(), dev.get_junction_x(), dev.get_bc(), dev.get_dex())      # for the phase replay
    totals0 = dev.get_totals()
    barrier()
    mark = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    dev.step(W, K)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    sampler.pause()                         # not during the phase replay / end-to-end runs (host-side interference)
    clocks = sampler.summary(mark)
    clocks["window"] = "timed loop only; clocks_whole_run covers warm-up + timed loop (+ the C5 run)"
    clocks_run = None
    totals = dev.get_totals() - totals0             # counters of exactly the K timed steps
    flags = int(dev.get_flags()[0])
    failed_step = dev.failed_step()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * nodes * K / (ms_total * 1e-3)

    # ---- the two per-step kernels, each timed alone: replay of the SAME steps from the saved state -----
    n_replay = min(K, 2000)
    (A0, q0), jx0, bc0, dex0 = saved
    dev.set_state(A0, q0)
    dev.set_junction_x(jx0)
    dev.set_bc(bc0)
    dev.set_dex(dex0)
    phase = phase_shares(dev, stream, W, n_replay, args.Nt, torch)

    # ---- ensemble statistics of the tree members: the one collective of this path -----
    from bloodflow_b200 import ensemble as en
    mean_p, coll_ms = None, None
    if dev.snapshot_count() > 0:
        with torch.cuda.stream(stream):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            stats = en.combine_moments(en.device_moments(dev, local_rank))
            c1.record(stream)
        torch.cuda.synchronize()
        coll_ms = c0.elapsed_time(c1)
        mean_p = float(stats['mean'].mean().item())

    # ---- end to end through the public API (rank-local, then max over ranks) -----------
    # one untimed warm-up repetition (first-touch costs of the allocator caches: cudaMalloc / cudaHostAlloc of the
    # store and the result arrays), then three complete timed repetitions (each: new network, upload, loop,
    # download); the median is reported, every repetition is listed
    e2e_warm = end_to_end(args, rank, local_rank, K)
    e2e_runs = [end_to_end(args, rank, local_rank, K) for _ in range(3)]
    e2e = sorted(e2e_runs, key=lambda r: r['seconds'])[1]
    e2e['seconds_all'] = [r['seconds'] for r in e2e_runs]
    e2e['seconds_warmup_rep'] = e2e_warm['seconds']
    te = torch.tensor([e2e['seconds']], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * nodes * K / float(te.item())
    an._dev.close()

    c5 = None
    if not args.no_c5:
        c5 = run_c5(args, rank, local_rank, world, torch, dist, sampler)
    sampler.stop()
    clocks_run = sampler.summary()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        peak_tf = fp64_peak(local_rank)
        ms_step = ms_total / K
        n_launches, n_dumps, n_art_launches = launches_in(schedule, W, K, args.Nt, mask)
        if schedule == _ffi.AF_SCHEDULE_PERSISTENT:
            # the artery kernel is persistent: a launch lives for its whole chunk of time steps (the junction
            # kernel runs beside it), so its duration per time step IS the step time
            art_us = ms_step * 1e3
            kernel_name = "k_artery_p"
            how = ("persistent kernel: CUDA-event time of the timed loop / time steps (each of the %d launches in the "
                   "timed region advances up to 64 time steps and is resident for all of them, waits for the junction "
                   "kernel beside it included)" % n_art_launches)
        else:
            art_us = phase['artery_share'] * ms_step * 1e3          # <= ms_per_step by construction
            kernel_name = "k_artery"
            how = ("share of the step from a replay of the timed steps with an event between the two kernels, times "
                   "the timed ms_per_step")
        its = totals[0] / float(dev.na * K)
        fl = flops_per_vertex_step(its)
        achieved_tf = fl * nodes / (art_us * 1e-6) / 1e12
        bytes_per_launch = nodes * 32.0
        hbm = bytes_per_launch / (art_us * 1e-6) / 1e9
        is_c4 = (args.order, args.Nx) == (ORDER, NX)
        cap = committed_capture("k_artery<256>") if is_c4 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args),
            "cycles_per_sec_per_network": 1.0 / (ms_step * 1e-3 * args.Nt),
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf,
                "traffic": (cap["dram_bytes_read"] + cap["dram_bytes_write"]) if cap else None,
                "peak_kind": "measured in this run: af_measure_fp64_peak (dependent-free DFMA chains, all SMs)",
                "kernel": kernel_name, "kernel_us": art_us, "kernel_us_per": "time step",
                "kernel_us_how": how,
                "schedule": {0: "per-step launches, grid-wide waits", 1: "per-step launches, dataflow flags",
                             2: "persistent kernels over chunks of <= 64 steps", 3: "fused"}[schedule],
                "per_step_kernels_timed_alone": {"k_boundary_us": phase['boundary_us_serialised'],
                                                 "k_artery_us": phase['artery_us_serialised'],
                                                 "note": "the per-step kernels (af_phase_boundary / af_phase_arteries) "
                                                         "replayed over the same steps with events between them: "
                                                         "how long each takes without any overlap"},
                "flops_per_vertex_step": fl, "newton_its_per_artery_step": its,
                "flops_model": "W(k) = (k+1)*110 + k*(140+50), k = measured mean linear solves per artery-step "
                               "(SURVEY 8(d)); redundant work of the parallel solver does not count",
                "note": "the order-8 tree's 8.2 MB state is L2-resident: FP64 is the binding roof, HBM is listed "
                        "beside it",
                "hbm": {"achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm / peaks["hbm_gbs"],
                        "peak_kind": peak_kind, "algorithmic_bytes_per_launch": bytes_per_launch,
                        "note": "32 B per vertex-step (read A,q; write A,q); untapered tree streams no coefficients"},
                "ncu": cap,
            },
            "phases": phase,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e['h2d'] / K,
                    "d2h_bytes_per_step": e2e['d2h'] / K, "seconds": float(te.item()),
                    "seconds_all_repeats_rank0": e2e['seconds_all'],
                    "seconds_untimed_warmup_repeat_rank0": e2e['seconds_warmup_rep'],
                    "api": "arteryfe.ArteryNetwork(param).solve(n_steps=K): create+upload, loop with %d-step "
                           "snapshot cadence, fetch of flow/area/pressure snapshots" % (args.Nt // NT_STORE)},
            "gpu_launches": n_launches,
            "clocks": clocks, "clocks_whole_run": clocks_run, "wall_seconds_host": wall,
            "status": {"flags": flags, "failed_step": failed_step, "artery_solves": int(totals[0]),
                       "junction_solves": int(totals[1]), "picard_its": int(totals[2]),
                       "mean_outlet_pressure": mean_p, "collective_ms": coll_ms},
        }
        if c5 is not None:
            line["ensemble_c5"] = c5
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.order, args.Nx, args.Nt, args.theta)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
