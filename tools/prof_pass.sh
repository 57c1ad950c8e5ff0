set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/p_plain.json 2> gpurun_out/p_plain.err && \
timeout 800 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_pass' -s 3 -c 1 -o gpurun_out/p_prof -f $CMD > gpurun_out/p_ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/p_ncu_full.log
