set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_full2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_full2.log; tail -25 gpurun_out/t_full2.log
timeout 600 python bench.py --workload glv256long --steps 2 --warmup 1 > gpurun_out/b256_long.json 2> gpurun_out/b256_long.err; tail -c 1200 gpurun_out/b256_long.json; tail -3 gpurun_out/b256_long.err
timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_pair2.json 2>&1; tail -c 300 gpurun_out/b256_pair2.json
