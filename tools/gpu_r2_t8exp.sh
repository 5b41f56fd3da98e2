# round 2: variants of the headline kernel (experiment builds selected with VA_ENGINE_LIB), parity check + device-resident rate
set -x
for v in "$@"; do
  lib=$PWD/vectorizedadjoint_b200/libva_exp_$v.so
  [ "$v" = base ] && lib=$PWD/vectorizedadjoint_b200/libva_engine.so
  VA_ENGINE_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "glv_batch_vs_oracle or half_norm or register_kernel_generations or full_size_properties" 2>&1 | tail -2
  for i in 1 2; do VA_ENGINE_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample --no-traffic-probe 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('T8EXP $v',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],4))"; done
done
