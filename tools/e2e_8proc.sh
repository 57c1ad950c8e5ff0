# the pipelined e2e probe in N independent processes at once, one per GPU: which host section grows with N?
N=${1:-8}
for k in $(seq 0 $((N-1))); do
  SST_DEVICE=$k python tools/e2e_async_probe.py > gpurun_out/e2e_${N}proc_$k.log 2>&1 &
done
wait
nproc; lscpu | grep -i "model name\|^CPU(s)\|NUMA\|Thread\|Socket" 
nvidia-smi topo -m | head -14
for k in 0 $((N-1)); do echo "== process $k"; grep "^depth 3 ce\|^depth 3 e\|^  \[\|per batch" gpurun_out/e2e_${N}proc_$k.log; done
