set -x
for v in a b c; do for i in 1 2; do
  VA_GLV_T8S=1 VA_ENGINE_LIB=$PWD/vectorizedadjoint_b200/libva_exp_$v.so timeout 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side --no-parity-sample --no-traffic-probe 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('T8SEXP $v',round(d['value']),d['ms_per_step'],round(d['roofline']['frac'],4),d['roofline']['kernel'][:12])"; done; done
