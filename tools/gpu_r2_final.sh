# round 2, second session: the whole -m gpu suite, smoke(), t8 vs t8s bit-identity, the contract line, the reference arm; optionally
# (SAN=1) compute-sanitizer over sanitize.py
set -x
O=gpurun_out/r02_final3
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 150 python tools/t8s_check.py > $O/t8s_check.log 2>&1; tail -2 $O/t8s_check.log
start=$(date +%s)
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench default took $(( $(date +%s) - start )) s"; tail -c 300 $O/bench_default.json
if [ "${REF:-0}" = 1 ]; then timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json; fi
if [ "${SAN:-0}" = 1 ]; then for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py > $O/san_$tool.log 2>&1; tail -2 $O/san_$tool.log
done; fi
