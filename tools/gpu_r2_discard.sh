# round 2: dead step blocks dropped from L2 (discard.global.L2) in the headline kernel -- parity, rate and DRAM traffic with / without
set -x
mkdir -p gpurun_out/r02_discard
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin_examples.py tests/test_gpu_torch_api.py -m gpu -q -x -k "glv or lotka" > gpurun_out/r02_discard/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_discard/pytest.log; tail -3 gpurun_out/r02_discard/pytest.log
timeout 120 python tools/t8s_check.py > gpurun_out/r02_discard/t8s_check.log 2>&1; tail -2 gpurun_out/r02_discard/t8s_check.log
for v in 0 1; do
  VA_T8_DISCARD=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample 2>gpurun_out/r02_discard/bench_$v.err | tee gpurun_out/r02_discard/bench_discard_$v.json | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);r=d['roofline'];print('DISCARD on=$v',round(d['value']),d['ms_per_step'],round(r['frac'],4),'traffic',r['traffic'],'alg',r['algorithmic_bytes'], r['traffic_source'][:40])"
done
