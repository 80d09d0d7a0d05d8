"""Executed instruction mix of one kernel from an ncu report (`--page source --csv`, SASS view):
thread-level counts of DFMA / DMUL / DADD / MUFU / shared-memory / barrier / other instructions and the
executed FP64 flop per vertex-step (2 per DFMA, 1 per DMUL / DADD) -- to be compared with the
algorithmic W(k) that bench.py's roofline uses.

usage: python scripts/ncu_opcode_mix.py report.ncu-rep <kernel regex> <vertices per launch> [launch index]"""
import csv
import io
import subprocess
import sys


def main(path, kernel, vertices, which=0):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel],
                         check=True, capture_output=True, text=True).stdout
    # the csv holds one block per profiled launch: "Kernel Name" line, header line, instruction lines
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == 'Kernel Name':
            cur = {'name': row[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None and row:
            cur['rows'].append(row)
    b = blocks[which]
    head, body = b['rows'][0], b['rows'][1:]
    col = {n: i for i, n in enumerate(head)}
    mix, warp_mix, samples = {}, {}, {}
    for r in body:
        op = r[col['Source']].split()
        if not op:
            continue
        op = op[1] if op[0].startswith('@') else op[0]
        base = op.split('.')[0]
        key = base if base in ('DFMA', 'DMUL', 'DADD', 'MUFU', 'LDS', 'STS', 'BAR', 'LDG', 'STG', 'SHFL', 'DSETP') else 'other'
        mix[key] = mix.get(key, 0) + int(r[col['Predicated-On Thread Instructions Executed']])
        warp_mix[key] = warp_mix.get(key, 0) + int(r[col['Instructions Executed']])
        samples[key] = samples.get(key, 0) + int(r[col['# Samples']])
    total = sum(mix.values())
    print('== %s (launch %d of %d in the report)' % (b['name'], which, len(blocks)))
    for k, v in sorted(mix.items(), key=lambda kv: -kv[1]):
        print('   %-6s %12d thread instr (%5.1f %%)  %10d warp instr  %6d stall samples' %
              (k, v, 100.0 * v / total, warp_mix[k], samples[k]))
    flop = 2 * mix.get('DFMA', 0) + mix.get('DMUL', 0) + mix.get('DADD', 0)
    print('   executed FP64 flop per vertex-step: %.0f  (DFMA %.0f, DMUL %.0f, DADD %.0f instructions per vertex)'
          % (flop / vertices, mix.get('DFMA', 0) / vertices, mix.get('DMUL', 0) / vertices, mix.get('DADD', 0) / vertices))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 0)
