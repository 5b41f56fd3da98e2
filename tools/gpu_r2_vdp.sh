set -x
mkdir -p gpurun_out/r02_vdp
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin_examples.py tests/test_gpu_tape.py -m gpu -q -x -k "vanderpol or harmonic or scalar or tape or recorded or device_buffers or backward or no_progress" 2>&1 | tail -5
for i in 1 2; do timeout 300 python bench.py --workload vdp --steps 10 --warmup 3 > gpurun_out/r02_vdp/b_vdp_$1_$i.json 2>gpurun_out/r02_vdp/err.txt; python -c "
import json;d=json.loads(open('gpurun_out/r02_vdp/b_vdp_$1_$i.json').read().strip().splitlines()[-1]);print('VDP',d['value'],d['ms_per_step'])"; done
timeout 300 python bench.py --workload harmonic --steps 5 --warmup 3 | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('HO',d['value'],d['ms_per_step'])"
