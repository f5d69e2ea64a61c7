#!/bin/bash
# A/B bench of alternative builds of liboqg (build_ab/*.so) on the GPU box:
#   tools/ab.sh [ruptures]   (run under gpurun)
R=${1:-30000}
for v in "" build_ab/*.so; do
  OQG_LIB=$v python bench.py --ruptures $R --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-28s %.4g evals/s  kernels ms %s' % ('$v' or 'default', d['value'], [round(x,2) for x in d['roofline']['kernel_ms']]))"
done
