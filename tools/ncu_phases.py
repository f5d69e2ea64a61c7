"""Per-phase shares of the fused kernel from an ncu report: warp instructions
and stall samples of the prologue, the rupture-tile fetch, the distance stage
(D), the rupture-only model stage (Q), the per-pair stage (P), the per-IMT
model evaluation (E), the PoE / level loop (L) and the flush of the partial
sums (F).

    OQG_LIB=<the profiled liboqg.so> python tools/ncu_phases.py REPORT.ncu-rep \
        <mangled-substring> <kernel-name-substring>

The SASS of the kernel is in program order; an instruction belongs to the
phase of the last instruction before it (itself included) whose source line
lies inside k_fused in oqg.cu, the phases being the line ranges between the
`// ----` markers of the kernel (inlined model / geometry / math code carries
the line of its own file and inherits the phase of the call site).  Inside E
and L the split is by the first level-loop line."""
import os
import re
import sys
import csv
import subprocess
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import line_table  # noqa

SRC = os.path.join(ROOT, 'oq_engine_b200', 'csrc', 'oqg.cu')


def phase_ranges():
    """[(first line, phase)] of k_fused in the current oqg.cu"""
    lines = open(SRC).read().splitlines()

    def find(text, start=0):
        for i in range(start, len(lines)):
            if text in lines[i]:
                return i + 1
        raise KeyError(text)
    k0 = find('k_fused(const FusedArgs A) {')
    marks = [
        (k0, 'prologue'),
        (find('for (int64_t r0 = r_begin', k0), 'fetch'),
        (find('// ---- distances', k0), 'D'),
        (find('// ---- rupture-only part', k0), 'Q'),
        (find("// s_rup is dead from here on", k0), 'fetch'),
        (find('// ---- the IMT-independent part', k0), 'P'),
        (find('// ---- GMPE + PoEs', k0), 'E'),
        (find('const double xa = -inv * A.xscale;', k0), 'L'),
        (find('if (FAST && (--tiles_left', k0), 'F'),
        (find('#undef DGET', k0), 'epilogue'),
    ]
    kend = find('// part [RC][G][L][N]', k0)
    return marks, k0, kend


def main(rep, func, kernel):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    pick = [i for i in starts if kernel in rows[i][1]][0]
    end = min([i for i in starts if i > pick] + [len(rows)])
    rows = rows[pick:end]
    hdr = rows[1]
    ii, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall_cols = [c for c in ('stall_long_sb', 'stall_short_sb', 'stall_wait',
                              'stall_barrier', 'stall_not_selected')
                  if c in hdr]
    sass = [r for r in rows[2:] if len(r) == len(hdr)]
    table = line_table(func)
    assert len(table) == len(sass), (len(table), len(sass))
    marks, k0, kend = phase_ranges()
    # the poe_table helpers sit above the kernel in oqg.cu: level loop
    poe0 = [i + 1 for i, ln in enumerate(open(SRC).read().splitlines())
            if 'sat_f32(double x64)' in ln][0]

    def phase_of_line(line):
        if poe0 <= line < k0:
            return 'L'
        if not k0 <= line < kend:
            return None
        cur = None
        for first, ph in marks:
            if line >= first:
                cur = ph
        return cur
    # the staged model code: the functions of every Fast<> in oqg_fast.cuh
    fast = open(os.path.join(ROOT, 'oq_engine_b200', 'csrc', 'oqg_fast.cuh')
                ).read().splitlines()
    fast_marks = []
    for i, ln in enumerate(fast):
        if '__device__' not in ln:
            continue
        for pat, ph in ((' S site(', 'prologue'), (' Q rup(', 'Q'),
                        (' P pair(', 'P'), (' void eval(', 'E')):
            if pat in ln:
                fast_marks.append((i + 1, ph))
                break
        else:
            # helpers shared by several stages (CY14 `one`, CB14 `base`):
            # their inlined copies inherit the phase of the call site
            if 'static' in ln:
                fast_marks.append((i + 1, None))

    def fast_phase(line):
        cur = None
        for first, ph in fast_marks:
            if line >= first:
                cur = ph
        return cur
    agg = collections.OrderedDict((ph, collections.Counter()) for ph in (
        'prologue', 'fetch', 'D', 'Q', 'P', 'E', 'L', 'F', 'epilogue'))
    cur, in_loop = 'prologue', False
    for r, (_off, (fname, line), _txt) in zip(sass, table):
        if fname == 'oqg.cu':
            ph = phase_of_line(line)
            # operands defined before the loop (coefficient rows, levels,
            # table address) are loaded where they are used: once inside
            # the loop code, a prologue line does not change the phase
            if ph and not (ph == 'prologue' and in_loop):
                cur = ph
                in_loop = in_loop or ph != 'prologue'
        elif fname == 'oqg_geo.cuh' and in_loop:
            cur = 'D'
        elif fname == 'oqg_fast.cuh':
            ph = fast_phase(line)
            if ph and (in_loop or ph != 'prologue'):
                cur = ph if not (ph == 'prologue' and in_loop) else cur
                in_loop = in_loop or ph != 'prologue'
        a = agg[cur]
        a['inst'] += float(r[ii] or 0)
        a['samp'] += float(r[isamp] or 0)
        for c in stall_cols:
            a[c] += float(r[hdr.index(c)] or 0)
    tot_i = sum(a['inst'] for a in agg.values())
    tot_s = sum(a['samp'] for a in agg.values())
    print('kernel', rows[0][1])
    print('total warp instructions %.4g, samples %d' % (tot_i, tot_s))
    print('| phase | instructions | samples (time) | ' + ' | '.join(stall_cols)
          + ' |')
    print('|---|---|---|' + '---|' * len(stall_cols))
    names = {'prologue': 'prologue (tables, site terms)',
             'fetch': 'rupture tile fetch + barriers',
             'D': 'D distances', 'Q': 'Q rupture-only model terms',
             'P': 'P per-pair model terms', 'E': 'E per-IMT model evaluation',
             'L': 'L PoE level loop', 'F': 'F flush of the partial sums',
             'epilogue': 'epilogue'}
    for ph, a in agg.items():
        print('| %s | %.1f %% | %.1f %% | %s |' % (
            names[ph], 100 * a['inst'] / tot_i, 100 * a['samp'] / tot_s,
            ' | '.join('%.1f %%' % (100 * a[c] / max(tot_s, 1))
                       for c in stall_cols)))


if __name__ == '__main__':
    main(*sys.argv[1:])
