# one GPU call: the whole -m gpu suite, smoke(), the contract bench line, the reference arm, compute-sanitizer on the new kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_full.log; tail -5 gpurun_out/t_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; tail -c 700 gpurun_out/bench_main.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 500 gpurun_out/bench_ref.json
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/san_$tool.log 2>&1; tail -3 gpurun_out/san_$tool.log
done
