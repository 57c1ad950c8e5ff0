set -x
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --large-factor 0"
$CMD > gpurun_out/r1f_plain.json 2> gpurun_out/r1f_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches.csv $CMD > gpurun_out/r1f_ncu_launch.log 2>&1
echo "launchlist rc=$?"
timeout 800 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_pass|k_build_table|k_transpose|k_classify|k_stage' -c 24 -o gpurun_out/r1f_prof -f $CMD > gpurun_out/r1f_ncu_full.log 2>&1
echo "full rc=$?"
