#!/bin/bash
# Final GPU pass of a round: tests, smoke, bench (both arms), launch list, full ncu capture and
# full-size traffic / pipe metrics of the fused kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_final.sh TAG'
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused -c 2 \
  -o gpurun_out/${TAG}_fused -f python bench.py --ruptures 8000 --steps 1 --warmup 0 --no-cpu > gpurun_out/${TAG}_ncu_f.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_fused -c 2 --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/${TAG}_traffic.log 2>&1
python tools/bench_post.py > gpurun_out/${TAG}_bench_post.json 2>/dev/null
python tools/bench_mesh.py > gpurun_out/${TAG}_bench_mesh.json 2>/dev/null
python tools/bench_sources.py > gpurun_out/${TAG}_bench_sources.json 2>/dev/null
cut -c1-260 gpurun_out/${TAG}_bench.json; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
