"""Per-kernel metrics of an ncu capture of bench.py as JSON, for the fields of
the bench line that only a profiler can measure (DRAM traffic, warp
instructions, pipe utilisation):

    python tools/ncu_metrics.py REPORT.ncu-rep FINGERPRINT.txt RUPTURES SITES \
        > profiles/r2_fused_metrics.json

FINGERPRINT.txt holds oq_engine_b200._build.fingerprint() of the build that
was profiled (written on the GPU box next to the report); bench.py uses the
numbers only when its own build has the same fingerprint.
"""
import sys
import csv
import json
import subprocess

FAMILY = {0: 'BooreEtAl2014', 1: 'ChiouYoungs2014', 2: 'AbrahamsonEtAl2014',
          3: 'CampbellBozorgnia2014', 4: 'BooreAtkinson2008',
          5: 'AtkinsonBoore2006', 6: 'AkkarEtAl2014'}
WANT = {
    'gpu__time_duration.sum': 'ncu_ms',
    'smsp__inst_executed.sum': 'warp_inst',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed':
        'lsu_data_pipe_pct',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active':
        'fp64_pipe_pct',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active':
        'fma_pipe_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active':
        'xu_pipe_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'launch__registers_per_thread': 'registers',
    'launch__block_size': 'block_size', 'launch__grid_size': 'grid_size',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum':
        'shared_bank_conflicts',
}
UNIT = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.,
        'ms': 1., 'us': 1e-3, 's': 1e3}


def main(rep, fpfile, U, N):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        if 'k_fused' not in name:
            continue
        fam = int(name.split('<')[1].split(',')[0].replace('(int)', ''))
        d = {'kernel': name}
        for col, key in WANT.items():
            if col in hdr:
                i = hdr.index(col)
                d[key] = float(r[i].replace(',', '')) * UNIT.get(units[i], 1.)
        d['dram_bytes'] = d.pop('dram_read') + d.pop('dram_write')
        d['warp_inst_per_pair'] = d['warp_inst'] / (float(U) * float(N))
        kernels[FAMILY[fam]] = d
    with open(fpfile) as f:
        fp = f.read().strip()
    json.dump({'report': rep.split('/')[-1], 'kernel_fingerprint': fp,
               'ruptures': int(U), 'sites': int(N), 'kernels': kernels},
              sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main(*sys.argv[1:])
