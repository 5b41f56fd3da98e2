set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; tail -c 300 gpurun_out/bench_main.json
timeout 600 python -m pytest tests -x -q -m gpu -k "glv" > gpurun_out/t_glv.log 2>&1; tail -3 gpurun_out/t_glv.log
