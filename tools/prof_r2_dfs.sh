set -x
# ncu --set full of the depth-first enumeration pass on the C4 batch (one B200)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_dfs' -s 3 -c 2 -o gpurun_out/r2_dfs -f python tools/quick_enum.py > gpurun_out/r2_dfs_ncu.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_dfs.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2_dfs_source.csv 2>/dev/null
ncu -i gpurun_out/r2_dfs.ncu-rep --page raw --csv > gpurun_out/r2_dfs_raw.csv 2>/dev/null
ls -la gpurun_out/
