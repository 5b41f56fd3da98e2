set -x
mkdir -p gpurun_out/r02_oct
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "quad or glv16 or glv_batch_vs_oracle or reference_example_data or synthetic_vs_reference or every_reference_tableau or half_norm or backward or finite_difference" 2>&1 | tail -15
for pol in auto store; do timeout 300 python bench.py --workload glv16 --steps 5 --warmup 3 --ckpt-policy $pol > gpurun_out/r02_oct/b16_$pol.json 2>gpurun_out/r02_oct/err_$pol.txt; python -c "
import json;d=json.loads(open('gpurun_out/r02_oct/b16_$pol.json').read().strip().splitlines()[-1]);print('GLV16 $pol',d['kernel'],d['value'],d['ms_per_step'],d['roofline']['frac'])"; done
