set -x
# Round-2 evidence: plain bench line, ncu launch list of the same command, ncu --set full of the step's kernels (one B200)
TAG=${1:-r2}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --large-factor 0"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launchlist rc=$?"
timeout 800 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_dfs|k_build_table|k_transpose|k_classify|k_stage' -c 24 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --print-source cuda,sass --csv -k regex:k_explain_dfs > gpurun_out/${TAG}_dfs_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof.ncu-rep
ls -la gpurun_out | tail -8
