#!/bin/bash
# gpu_profile.sh <tag> [keep-list] — one ncu --set full capture of the render kernel per BASELINE workload (run under gpurun,
# 1 GPU).  The reports are reduced ON THE BOX (gpurun merges at most 64 MiB back): tests/ncu_extract.py writes
# profiles/<tag>_<workload>_<mode>.raw.csv + profiles/r02_kernel_counters.json, copies land in gpurun_out/<tag>_extract/;
# only the reports named in keep-list (default "cfg3_voxel") come back whole for tests/ncu_lines.py.
tag=${1:-r02}
keep=${2:-cfg3_voxel}
mkdir -p gpurun_out/${tag}_extract
args=""
for wm in cfg1:trilinear cfg2:levelset cfg3:voxel cfg4:deep cfg4:deepshadow; do
  w=${wm%%:*}; m=${wm##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gx_render -s 1 -c 1 -f -o gpurun_out/${tag}_${w}_${m} \
      python tests/prof_frame.py --workload $w --mode $m --frames 2 > gpurun_out/${tag}_${w}_${m}.log 2>&1
  tail -1 gpurun_out/${tag}_${w}_${m}.log
  args="$args $w:$m:tex=gpurun_out/${tag}_${w}_${m}.ncu-rep"
done
python tests/ncu_extract.py $args > gpurun_out/${tag}_extract/extract.log 2>&1
cp profiles/${tag}_*.raw.csv profiles/r02_kernel_counters.json gpurun_out/${tag}_extract/ 2>/dev/null
for f in gpurun_out/${tag}_*.ncu-rep; do
  b=$(basename $f .ncu-rep); b=${b#${tag}_}
  case " $keep " in *" $b "*) ;; *) rm -f $f ;; esac
done
du -sh gpurun_out
