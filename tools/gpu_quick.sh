#!/bin/bash
# Quick GPU pass: tests, smoke, both bench arms, secondary benches (no ncu).
#   gpurun --timeout 1800 -- 'bash tools/gpu_quick.sh TAG'
TAG=${1:-quick}
mkdir -p gpurun_out
python -c "from oq_engine_b200 import _build; print(_build.kernel_fingerprint())" > gpurun_out/${TAG}_fingerprint.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
python tools/bench_post.py > gpurun_out/${TAG}_bench_post.json 2>/dev/null
python tools/bench_mesh.py > gpurun_out/${TAG}_bench_mesh.json 2>/dev/null
python tools/bench_sources.py > gpurun_out/${TAG}_bench_sources.json 2>/dev/null
cut -c1-400 gpurun_out/${TAG}_bench.json; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
cat gpurun_out/${TAG}_bench_post.json | cut -c1-600
