#!/bin/bash
# GPU pass of round 2: tests, smoke, bench (both arms), launch list, full ncu capture
# (--import-source) and full-size per-launch metrics of the fused kernels + the build fingerprint
# the metrics belong to.
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2.sh TAG'
# (gpurun copies back at most 64 MiB: the post-processing capture is a call of its own,
#  tools/gpu_post.sh)
TAG=${1:-r2}
mkdir -p gpurun_out
python -c "from oq_engine_b200 import _build; print(_build.kernel_fingerprint())" > gpurun_out/${TAG}_fingerprint.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-strong > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fused -c 2 \
  -o gpurun_out/${TAG}_fused -f python bench.py --ruptures 8000 --steps 1 --warmup 0 --no-cpu --no-strong > gpurun_out/${TAG}_ncu_f.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__block_size,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:k_fused -c 2 -o gpurun_out/${TAG}_traffic -f python bench.py --steps 1 --warmup 0 --no-cpu --no-strong > gpurun_out/${TAG}_traffic.log 2>&1
python tools/bench_post.py > gpurun_out/${TAG}_bench_post.json 2>/dev/null
python tools/bench_mesh.py > gpurun_out/${TAG}_bench_mesh.json 2>/dev/null
python tools/bench_sources.py > gpurun_out/${TAG}_bench_sources.json 2>/dev/null
python tools/bench_post.py --sites 10000 --rlzs 1000 > gpurun_out/${TAG}_bench_post_r1000.json 2>/dev/null
python tools/bench_host_stream.py > gpurun_out/${TAG}_host_stream.json 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_postproc.py tests/test_gmf.py -m gpu -x -q > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/${TAG}_memcheck.log
du -sh gpurun_out; ls -la gpurun_out | sort -k5 -n | tail -4
cut -c1-260 gpurun_out/${TAG}_bench.json; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
