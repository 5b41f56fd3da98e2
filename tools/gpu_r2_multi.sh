# round 2, multi-GPU call (gpurun --gpus N): host topology, the N>1 parity tests, the contract line both ways (one rank per GPU
# under torchrun; one process with a multi-device engine), host->device copy ceilings per host-buffer kind
set -x
N=${1:-2}
O=gpurun_out/r02_multi_${N}
mkdir -p $O
{ nvidia-smi -L; nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread"; ls /sys/devices/system/node/; free -g | head -2; \
  for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class)"; fi; done; } > $O/topology.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.txt 2>&1; tail -15 $O/pytest_multi.txt
timeout 120 vectorizedadjoint_b200/examples/build/multi_lotka $N > $O/multi_lotka.txt 2>&1; cat $O/multi_lotka.txt
timeout 300 python - > $O/h2d.txt 2>&1 <<PY
import vectorizedadjoint_b200 as va, json
devs=list(range($N))
for flags,name in ((0,"pinned"),(1,"write_combined"),(2,"numa_local")):
    for sel in ([0],devs):
        per,agg=va.measure_h2d_copy(sel,nbytes=2<<30,reps=3,flags=flags)
        print(json.dumps({"flags":name,"devices":sel,"per_gpu_gbs":per,"aggregate_gbs":agg}),flush=True)
PY
cat $O/h2d.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_ranks.json 2> $O/bench_ranks.err; tail -c 1500 $O/bench_ranks.json; tail -3 $O/bench_ranks.err
timeout 900 python bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_devices.json 2> $O/bench_devices.err; tail -c 1500 $O/bench_devices.json; tail -3 $O/bench_devices.err
VA_BENCH_HOST_FLAGS=1 timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-check > $O/bench_ranks_wc.json 2> $O/bench_ranks_wc.err; tail -c 900 $O/bench_ranks_wc.json
