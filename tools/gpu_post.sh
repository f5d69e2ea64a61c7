#!/bin/bash
# ncu captures of the post-processing kernels (gpurun copies back at most 64 MiB: one launch with
# the source page, the others with the summary sections only)
#   gpurun --timeout 900 -- 'bash tools/gpu_post.sh TAG'
TAG=${1:-post}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_postclassical_reg" --launch-skip 2 -c 1 \
  -o gpurun_out/${TAG}_post8 -f python tools/bench_post.py --steps 1 > gpurun_out/${TAG}_ncu_post8.log 2>&1
timeout 300 ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats \
  --section MemoryWorkloadAnalysis --clock-control none -k regex:"k_postclassical_reg|k_hazard_maps|k_mean_rates" \
  --launch-skip 4 -c 6 -o gpurun_out/${TAG}_post -f python tools/bench_post.py --steps 1 > gpurun_out/${TAG}_ncu_post.log 2>&1
du -sh gpurun_out
