"""Condense an `ncu --set full` report into the few lines kept under profiles/.

usage: python scripts/ncu_summary.py report.ncu-rep [vertices per launch] > profiles/rNN_..._summary.txt
(reads the report with `ncu -i ... --page raw --csv`; no GPU needed).  With the number of vertices a
launch advances, the executed FP64 flop per vertex-step (2 per DFMA, 1 per DMUL / DADD) is printed next to
the algorithmic W(k) of bench.py."""
import csv
import io
import subprocess
import sys

KEEP = (
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic',
    # L2 and DRAM throughput (north_star: "achieved HBM and L2 GB/s")
    'lts__t_bytes.sum', 'lts__t_bytes.sum.per_second', 'lts__t_sectors.sum', 'lts__t_sectors.sum.per_second',
    'lts__t_sector_hit_rate.pct',
    'lts__t_sectors.avg.pct_of_peak_sustained_elapsed', 'dram__bytes.sum.per_second',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    # executed FP64 work
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__inst_executed_pipe_fp64.sum',
)
STALL = 'smsp__average_warps_issue_stalled_'


def main(path, vertices=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {name: k for k, name in enumerate(head)}
    for r in body:
        print('== %s' % r[col['Kernel Name']])
        for m in KEEP:
            if m in col:
                print('   %s %s %s' % (m, r[col[m]], units[col[m]]))
        stalls = []
        for name, k in col.items():
            if name.startswith(STALL) and name.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(r[k]), name[len(STALL):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        try:
            # achieved L2 throughput: 32-byte sectors per second (the report's unit prefix is in the units row)
            k = col['lts__t_sectors.sum.per_second']
            val, unit = float(r[k].replace(',', '')), units[k]
            scale = {'sector/s': 1.0, 'Ksector/s': 1e3, 'Msector/s': 1e6, 'Gsector/s': 1e9, 'Tsector/s': 1e12,
                     'sector/ns': 1e9, 'sector/us': 1e6, 'sector/ms': 1e3}.get(unit)
            if scale:
                print('   achieved L2 throughput (lts__t_sectors.sum.per_second x 32 B): %.1f GB/s' % (val * scale * 32 / 1e9))
        except (KeyError, ValueError):
            pass
        if vertices:
            try:
                g = lambda m: float(r[col['smsp__sass_thread_inst_executed_op_%s_pred_on.sum' % m]].replace(',', ''))   # noqa: E731
                flop = 2 * g('dfma') + g('dmul') + g('dadd')
                print('   executed FP64 flop per vertex-step: %.0f (DFMA %.0f, DMUL %.0f, DADD %.0f per vertex)'
                      % (flop / vertices, g('dfma') / vertices, g('dmul') / vertices, g('dadd') / vertices))
            except (KeyError, ValueError):
                pass
        stalls.sort(reverse=True)
        print('   stalls: ' + ', '.join('%s=%.2f' % (n, v) for v, n in stalls[:7]))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0)
