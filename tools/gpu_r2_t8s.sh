# round 2: the warp-specialised headline kernel (va_glv_t8s.cu, VA_GLV_T8S=1) -- bit-identity with va_glv_t8.cu, then the rate.
# Every step under its own short timeout: a protocol error in the kernel is a hang, not a wrong number.
set -x
mkdir -p gpurun_out/r02_t8s
timeout 150 python tools/t8s_check.py > gpurun_out/r02_t8s/check.log 2>&1; rc=$?; echo "check rc=$rc" >> gpurun_out/r02_t8s/check.log; tail -12 gpurun_out/r02_t8s/check.log
[ $rc -ne 0 ] && exit 1
for v in ${T8S_VARIANTS:-1 0}; do for i in 1 2; do
  VA_GLV_T8S=$v timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample --no-traffic-probe 2>gpurun_out/r02_t8s/bench_$v.err | tee gpurun_out/r02_t8s/bench_${v}_$i.json | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('T8S $v',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],4),d['roofline']['kernel'][:12])"; done; done
