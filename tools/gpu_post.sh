#!/bin/bash
# ncu --set full capture of the post-processing kernels (grouped and one group per realization)
#   gpurun --timeout 900 -- 'bash tools/gpu_post.sh TAG'
TAG=${1:-post}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_postclassical_reg" --launch-skip 2 -c 1 \
  -o gpurun_out/${TAG}_post8 -f python tools/bench_post.py --steps 1 > gpurun_out/${TAG}_ncu_post8.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_postclassical_reg|k_hazard_maps" --launch-skip 4 -c 4 \
  -o gpurun_out/${TAG}_post -f python tools/bench_post.py --steps 1 > gpurun_out/${TAG}_ncu_post.log 2>&1
du -sh gpurun_out
