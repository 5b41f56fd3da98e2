# round 2, second session: the whole -m gpu suite, smoke(), compute-sanitizer over the new kernel paths, the contract line, the reference arm
set -x
mkdir -p gpurun_out/r02_final2
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final2/pytest_gpu.log; tail -4 gpurun_out/r02_final2/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final2/smoke.log 2>&1; tail -2 gpurun_out/r02_final2/smoke.log
for tool in memcheck racecheck synccheck initcheck; do
  VA_SANITIZE_ONLY=t8s timeout 600 compute-sanitizer --tool $tool python tools/sanitize.py > gpurun_out/r02_final2/san_t8s_$tool.log 2>&1; tail -3 gpurun_out/r02_final2/san_t8s_$tool.log
done
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r02_final2/bench_default.json 2> gpurun_out/r02_final2/bench_default.err; tail -c 600 gpurun_out/r02_final2/bench_default.json; grep -i "elapsed" gpurun_out/r02_final2/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_final2/bench_ref.json 2> gpurun_out/r02_final2/bench_ref.err; tail -c 400 gpurun_out/r02_final2/bench_ref.json
