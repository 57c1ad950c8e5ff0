set -x
# ncu --set full of the direct enumeration pass on the C4 batch tiled 16x (throughput regime)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_direct' -s 3 -c 1 -o gpurun_out/r2_direct16 -f python tools/large_phases.py > gpurun_out/r2_direct16_ncu.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_direct16.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_direct16_source.csv 2>/dev/null
ncu -i gpurun_out/r2_direct16.ncu-rep --page raw --csv > gpurun_out/r2_direct16_raw.csv 2>/dev/null
