#!/bin/bash
# A/B bench of alternative builds of liboqg (build_ab/*.so) on the GPU box:
#   tools/ab.sh [ruptures] [workload]   (run under gpurun)
R=${1:-30000}
W=${2:-cfg3}
for v in "" build_ab/*.so; do
  OQG_LIB=$v python bench.py --workload $W --ruptures $R --steps 2 --warmup 3 --no-cpu --no-strong 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-28s %.4g evals/s  kernels ms %s  executed %.3g' % ('$v' or 'default', d['value'], [round(x,2) for x in d['roofline']['kernel_ms']], (d['roofline'].get('executed') or {}).get('evals_per_step', 0)))"
done
