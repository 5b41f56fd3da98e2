# round 2, second session, gpurun --gpus N: the N>1 parity tests, the contract line under torchrun (driver's launch), the reference arm under torchrun
set -x
N=${1:-2}
O=gpurun_out/r02_final2_multi$N
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.txt 2>&1; tail -3 $O/pytest_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_ranks.json 2> $O/bench_ranks.err; tail -c 1200 $O/bench_ranks.json; tail -3 $O/bench_ranks.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N --impl reference --steps 1 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json
