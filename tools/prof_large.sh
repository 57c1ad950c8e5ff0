set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --large-factor 16"
$CMD > gpurun_out/l_plain.json 2> gpurun_out/l_plain.err && \
timeout 800 ncu --set full --clock-control none --import-source on -k 'regex:k_explain_pass' -s 12 -c 1 -o gpurun_out/l_prof -f $CMD > gpurun_out/l_ncu_full.log 2>&1
echo "full rc=$?"
