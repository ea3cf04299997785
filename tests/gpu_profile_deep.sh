tag=r02fin
mkdir -p gpurun_out/${tag}_extract
args=""
for wm in cfg4:deep cfg4:deepshadow; do
  w=${wm%%:*}; m=${wm##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gx_render -s 1 -c 1 -f -o gpurun_out/${tag}_${w}_${m} python tests/prof_frame.py --workload $w --mode $m --frames 2 > gpurun_out/${tag}_${w}_${m}.log 2>&1
  args="$args $w:$m:tex=gpurun_out/${tag}_${w}_${m}.ncu-rep"
done
python tests/ncu_extract.py $args > gpurun_out/${tag}_extract/extract.log 2>&1
cp profiles/${tag}_cfg4_*.raw.csv profiles/r02_kernel_counters.json gpurun_out/${tag}_extract/
rm -f gpurun_out/${tag}_cfg4_deep.ncu-rep
du -sh gpurun_out
