# A/B of the headline kernel: the default build against experiment builds libva_exp_<name>.so (VA_ENGINE_LIB), parity subset first.
# usage: gpurun -- 'bash tools/gpu_r2_ab.sh <tag> <name> [<name> ...]'
set -x
tag=$1; shift
mkdir -p gpurun_out/r02_$tag
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "glv" > gpurun_out/r02_$tag/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_$tag/pytest.log; tail -3 gpurun_out/r02_$tag/pytest.log
for rep in 1 2; do for v in default "$@"; do
  lib=$PWD/vectorizedadjoint_b200/libva_engine.so; [ $v != default ] && lib=$PWD/vectorizedadjoint_b200/libva_exp_$v.so
  VA_ENGINE_LIB=$lib timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample --no-traffic-probe 2>/dev/null | tee gpurun_out/r02_$tag/bench_${v}_$rep.json | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('AB $tag $v',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],4))"; done; done
