"""Per-source-line instruction counts and stall samples of one profiled
kernel: joins the SASS page of an ncu report with the line table of the
cubin inside liboqg.so (nvdisasm -g; needs -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <mangled-substring> [top]
        [stall-column|''] [kernel-name-substring]
"""
import os
import re
import sys
import csv
import subprocess
import tempfile
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(func_substr):
    tmp = tempfile.mkdtemp()
    # the library the report was taken with: OQG_LIB or the in-tree build
    subprocess.run(['cuobjdump', '-xelf', 'all',
                    os.path.abspath(os.environ.get('OQG_LIB') or os.path.join(
                        ROOT, 'oq_engine_b200', 'liboqg.so'))],
                   cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)],
                         capture_output=True, text=True).stdout
    table, cur, on = [], ('?', 0), False
    for ln in out.splitlines():
        if ln.startswith('.text.'):
            on = func_substr in ln
            continue
        if not on:
            continue
        mo = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mo:
            cur = (os.path.basename(mo.group(1)), int(mo.group(2)))
            continue
        mo = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
        if mo:
            table.append((int(mo.group(1), 16), cur, mo.group(2).strip()))
    return table


def main(rep, func, top=40, stall=None, kernel=None):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # one section per (kernel, view): keep the first section of the kernel
    # whose name holds `kernel` (default: the first)
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    pick = [i for i in starts if kernel is None or kernel in rows[i][1]][0]
    end = min([i for i in starts if i > pick] + [len(rows)])
    rows = rows[pick:end]
    hdr = rows[1]
    ia, ii, isamp = (hdr.index('Address'), hdr.index('Instructions Executed'),
                     hdr.index(stall or '# Samples'))
    sass = [(int(r[ia], 16), float(r[ii] or 0), float(r[isamp] or 0), r[1])
            for r in rows[2:] if len(r) == len(hdr)]
    table = line_table(func)
    assert len(table) == len(sass), (len(table), len(sass))
    base = sass[0][0]
    agg = collections.defaultdict(lambda: [0., 0.])
    pipes = collections.Counter()
    for (addr, inst, samp, txt), (off, where, txt2) in zip(sass, table):
        assert addr - base == off, (addr - base, off)
        agg[where][0] += inst
        agg[where][1] += samp
        op = txt.split()[0] if not txt.startswith('@') else txt.split()[1]
        pipes[op.split('.')[0]] += inst
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    print('total warp instructions %.4g, samples %d' % (tot_i, tot_s))
    src = {}
    for (f, line), (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0]
                                          )[:int(top)]:
        if f not in src:
            try:
                src[f] = open(os.path.join(ROOT, 'oq_engine_b200', 'csrc', f)
                              ).read().splitlines()
            except OSError:
                src[f] = []
        text = src[f][line - 1].strip()[:80] if 0 < line <= len(src[f]) else ''
        print('%5.1f%% inst %5.1f%% samp  %s:%d  %s' % (
            100 * inst / tot_i, 100 * samp / tot_s, f, line, text))
    print('opcodes:', ', '.join('%s %.1f%%' % (k, 100 * v / tot_i)
                                for k, v in pipes.most_common(24)))


if __name__ == '__main__':
    main(*sys.argv[1:])
