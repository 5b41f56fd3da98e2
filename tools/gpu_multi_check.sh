# multi-GPU call (gpurun --gpus N): the contract bench line, the N = 256 workload and the reference arm under torchrun
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 600 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
timeout 600 $TR --master-port 29512 bench.py --gpus $N --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_${N}gpu.json 2> gpurun_out/b256_${N}gpu.err; tail -c 400 gpurun_out/b256_${N}gpu.json; tail -3 gpurun_out/b256_${N}gpu.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; tail -c 300 gpurun_out/bench_ref_${N}gpu.json
