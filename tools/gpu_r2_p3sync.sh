# round 2 (negative result, the variant is no longer in the source): phase 3 of the headline kernel without its per-step 64-thread barrier
# of the default build against the VA_T8_P3SYNC=0 build (libva_exp_p3sync0.so)
set -x
mkdir -p gpurun_out/r02_p3sync
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "glv" > gpurun_out/r02_p3sync/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_p3sync/pytest.log; tail -3 gpurun_out/r02_p3sync/pytest.log
for v in new old new old; do
  lib=$PWD/vectorizedadjoint_b200/libva_engine.so; [ $v = old ] && lib=$PWD/vectorizedadjoint_b200/libva_exp_p3sync0.so
  VA_ENGINE_LIB=$lib timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample --no-traffic-probe 2>/dev/null | tee gpurun_out/r02_p3sync/bench_$v.json | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('P3SYNC $v',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],4))"; done
