set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_full3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_full3.log; tail -12 gpurun_out/t_full3.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/san_memcheck.log 2>&1; tail -3 gpurun_out/san_memcheck.log
timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_pair3.json 2>&1; tail -c 200 gpurun_out/b256_pair3.json
