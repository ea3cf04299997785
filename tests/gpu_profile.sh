#!/bin/bash
# gpu_profile.sh <tag> — one ncu --set full capture of the render kernel per BASELINE workload (run under gpurun, 1 GPU).
# Reports land in gpurun_out/<tag>_<workload>.ncu-rep; tests/ncu_extract.py turns them into profiles/*.raw.csv here.
tag=${1:-r02}
mkdir -p gpurun_out
for wm in cfg1:trilinear cfg2:levelset cfg3:voxel cfg4:deep cfg4:deepshadow; do
  w=${wm%%:*}; m=${wm##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gx_render -s 1 -c 1 -f -o gpurun_out/${tag}_${w}_${m} \
      python tests/prof_frame.py --workload $w --mode $m --frames 2 > gpurun_out/${tag}_${w}_${m}.log 2>&1
  tail -1 gpurun_out/${tag}_${w}_${m}.log
done
