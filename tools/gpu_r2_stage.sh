set -x
mkdir -p gpurun_out/r02_stage
timeout 900 python bench.py > gpurun_out/r02_stage/bench_default.json 2> gpurun_out/r02_stage/bench_default.err; tail -c 300 gpurun_out/r02_stage/bench_default.json
VA_T8_NO_STAGE=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample > gpurun_out/r02_stage/bench_nostage.json 2>/dev/null; tail -c 200 gpurun_out/r02_stage/bench_nostage.json
