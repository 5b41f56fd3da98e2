set -x
mkdir -p gpurun_out/r02_final2
timeout 120 python tools/t8s_check.py > gpurun_out/r02_final2/t8s_check.log 2>&1; tail -2 gpurun_out/r02_final2/t8s_check.log
VA_SANITIZE_ONLY=t8s timeout 600 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/r02_final2/san_t8s_racecheck.log 2>&1; tail -3 gpurun_out/r02_final2/san_t8s_racecheck.log
VA_GLV_T8S=1 timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample --no-traffic-probe 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('T8S',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],4))"
start=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02_final2/bench_default.json 2> gpurun_out/r02_final2/bench_default.err; echo "bench default took $(( $(date +%s) - start )) s"; tail -c 600 gpurun_out/r02_final2/bench_default.json
