"""Stall samples and executed instructions of k_artery by PHASE of the kernel body.

The ncu source page (SASS view, `--page source --csv`) gives samples per instruction address; nvdisasm -gi on
the cubin of the same build gives, per instruction offset, the chain of inlined source lines.  The outermost
line inside the kernel body selects the phase (line ranges of the kernel given on the command line as
name:first-last,...).

usage: python scripts/ncu_phase_profile.py report.ncu-rep disasm.sass <mangled kernel> <kernel regex> name:lo-hi[,name:lo-hi...]"""
import collections
import csv
import io
import re
import subprocess
import sys


def line_map(sass_path, mangled, src_name='af_kernels.cu'):
    """offset -> outermost line of src_name in the inline chain (the line in the kernel body)."""
    out, lines, inside = {}, [], False
    pat_line = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    pat_ins = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/')
    for ln in open(sass_path):
        if ln.startswith('.text.' + mangled + ':'):
            inside = True
            continue
        if inside and ln.startswith('//---------------------'):
            break
        if not inside:
            continue
        m = pat_line.search(ln)
        if m:
            lines.append(m)
            continue
        m = pat_ins.match(ln)
        if m:
            body = None
            # the chain is printed innermost first; the last entry naming src_name is the outermost
            for c in lines:
                if c.group(3):
                    if c.group(3).endswith(src_name):
                        body = int(c.group(4))
                elif c.group(1).endswith(src_name):
                    body = int(c.group(2))
            if body is not None:
                last = body
            out[int(m.group(1), 16)] = body if body is not None else out.get('last')
            if body is not None:
                out['last'] = body
            lines = []
    out.pop('last', None)
    return out


def main(report, sass, mangled, kernel, phases):
    ph = []
    for item in phases.split(','):
        name, rng = item.split(':')
        lo, hi = rng.split('-')
        ph.append((name, int(lo), int(hi)))
    lm = line_map(sass, mangled)
    txt = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel],
                         check=True, capture_output=True, text=True).stdout
    head, base = None, None
    samples, instr, fp64 = collections.Counter(), collections.Counter(), collections.Counter()
    n_blocks = 0
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == 'Kernel Name':
            n_blocks += 1
            if n_blocks > 1:
                break
            continue
        if row and row[0] == 'Address':
            head = {n: i for i, n in enumerate(row)}
            continue
        if not row or head is None:
            continue
        addr = int(row[0], 16)
        if base is None:
            base = addr
        line = lm.get(addr - base)
        name = 'unmapped'
        if line is not None:
            name = 'other'
            for pname, lo, hi in ph:
                if lo <= line <= hi:
                    name = pname
                    break
        samples[name] += int(row[head['# Samples']])
        n = int(row[head['Instructions Executed']])
        instr[name] += n
        op = row[head['Source']].split()
        op = [t for t in op if not t.startswith('@')]
        if op and op[0].split('.')[0] in ('DFMA', 'DMUL', 'DADD'):
            fp64[name] += n
    ts, ti = sum(samples.values()), sum(instr.values())
    print('phase                      samples      %%   warp-instr      %%   fp64-instr')
    for name in [p[0] for p in ph] + ['other', 'unmapped']:
        if instr[name] or samples[name]:
            print('%-24s %8d  %5.1f  %11d  %5.1f  %11d' % (name, samples[name], 100.0 * samples[name] / max(ts, 1),
                                                           instr[name], 100.0 * instr[name] / max(ti, 1), fp64[name]))


if __name__ == '__main__':
    main(*sys.argv[1:6])
