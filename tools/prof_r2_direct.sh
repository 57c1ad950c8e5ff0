set -x
# ncu --set full of the direct enumeration pass on the C4 batch (one B200)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_direct' -s 3 -c 2 -o gpurun_out/r2_direct -f python tools/quick_enum.py > gpurun_out/r2_direct_ncu.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_direct.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_direct_source.csv 2>/dev/null
ncu -i gpurun_out/r2_direct.ncu-rep --page raw --csv > gpurun_out/r2_direct_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
