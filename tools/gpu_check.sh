#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list and full capture of the fused kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused -c 2 \
  -o gpurun_out/${TAG}_fused -f python bench.py --ruptures 8000 --steps 1 --warmup 0 --no-cpu > gpurun_out/${TAG}_ncu_f.log 2>&1
ls -la gpurun_out
