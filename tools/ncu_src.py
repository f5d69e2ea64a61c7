"""Per-source-line and per-phase tables of one profiled kernel from an ncu
report alone (captured with --set full --import-source on; the library is
compiled with -lineinfo):

    python tools/ncu_src.py REPORT.ncu-rep KERNEL-SUBSTRING [top]

`ncu --page source --print-source cuda,sass` lists every source line followed
by its SASS instructions, so the line table of the cubin is not needed (the
older tools/ncu_lines.py joins with nvdisasm output of the CURRENT
liboqg.so, which must then be the profiled build).

Phases: the SASS of the fused kernel is cut at its barriers (BAR.SYNC) in
address order: prologue / tile wait / distances + rupture stage / pair stage
/ evaluation (model + PoEs + accumulation).
"""
import os
import re
import sys
import csv
import subprocess
import collections


def load(rep, kernel):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv',
                          '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    insts = {}   # address -> (file, line, sass, inst, samples, dict)
    fname, func, hdr, cur = None, None, None, None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            fname = os.path.basename(r[1])
        elif len(r) == 2 and r[0] == 'Function Name':
            func = r[1]
        elif r and r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) == len(hdr) and kernel in (func or ''):
            if r[0]:
                cur = (fname, int(r[0]), r[1].strip())
            elif r[2].startswith('0x'):
                d = dict(zip(hdr[4:], r[4:]))
                insts[int(r[2], 16)] = (cur, r[3].strip(), d)
    return insts


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.


def main(rep, kernel, top=40):
    insts = load(rep, kernel)
    addrs = sorted(insts)
    tot_i = sum(num(insts[a][2]['Instructions Executed']) for a in addrs)
    tot_s = sum(num(insts[a][2]['# Samples']) for a in addrs)
    print('%s: %d SASS instructions, %.4g warp instructions executed, '
          '%d samples' % (kernel, len(addrs), tot_i, tot_s))
    # ---- phases
    seg, agg = 0, collections.defaultdict(lambda: [0., 0., 0, collections.Counter()])
    for a in addrs:
        cur, sass, d = insts[a]
        i, s = num(d['Instructions Executed']), num(d['# Samples'])
        g = agg[seg]
        g[0] += i
        g[1] += s
        g[2] += 1
        op = sass.split()[1] if sass.startswith('@') else sass.split()[0]
        g[3][op.split('.')[0]] += i
        if 'BAR.SYNC' in sass:
            seg += 1
    print('\n| phase (between barriers) | SASS | instructions | samples | top opcodes (% of all instructions) |')
    print('|---|---|---|---|---|')
    for k, g in agg.items():
        print('| %d | %d | %.1f %% | %.1f %% | %s |' % (
            k, g[2], 100 * g[0] / tot_i, 100 * g[1] / tot_s, ', '.join(
                '%s %.1f' % (o, 100 * v / tot_i) for o, v in g[3].most_common(6))))
    # ---- lines
    lines = collections.defaultdict(lambda: [0., 0.])
    stalls = collections.Counter()
    ops = collections.Counter()
    for a in addrs:
        cur, sass, d = insts[a]
        i, s = num(d['Instructions Executed']), num(d['# Samples'])
        lines[cur][0] += i
        lines[cur][1] += s
        op = sass.split()[1] if sass.startswith('@') else sass.split()[0]
        ops[op.split('.')[0]] += i
        for k, v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k:
                stalls[k] += num(v)
    print('\n| instructions | samples | line | source |')
    print('|---|---|---|---|')
    for cur, (i, s) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        print('| %.1f %% | %.1f %% | %s:%d | `%s` |' % (
            100 * i / tot_i, 100 * s / tot_s, cur[0], cur[1], cur[2][:70]))
    print('\nopcodes: ' + ', '.join('%s %.1f %%' % (o, 100 * v / tot_i)
                                    for o, v in ops.most_common(20)))
    ts = sum(stalls.values())
    print('stall samples: ' + ', '.join('%s %.1f %%' % (k[6:], 100 * v / ts)
                                        for k, v in stalls.most_common(10)))


if __name__ == '__main__':
    main(*sys.argv[1:])
