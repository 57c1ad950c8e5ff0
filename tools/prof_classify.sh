set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --large-factor 0"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_classify' -c 2 -o gpurun_out/cls_prof -f $CMD > gpurun_out/cls_ncu.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/cls_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/cls_source.csv 2>/dev/null
ncu -i gpurun_out/cls_prof.ncu-rep --page source --csv > gpurun_out/cls_sass.csv 2>/dev/null
ncu -i gpurun_out/cls_prof.ncu-rep --page raw --csv > gpurun_out/cls_raw.csv 2>/dev/null
ls -la gpurun_out/
