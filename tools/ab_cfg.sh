#!/bin/bash
# A/B of the fused kernel's CTA shapes on the GPU box (run under gpurun):
#   tools/ab_cfg.sh [ruptures] [workload]
R=${1:-30000}
W=${2:-cfg3}
for cfg in "--plain-table" "--fused-subs 2" "--fused-subs 3" "--fused-subs 4"; do
  python bench.py --workload $W --ruptures $R --steps 2 --warmup 3 --no-cpu $cfg 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-18s %.4g evals/s  kernels ms %s' % ('$cfg', d['value'], [round(x,2) for x in d['roofline']['kernel_ms']]))"
done
