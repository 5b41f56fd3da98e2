python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_r1_final.err
for w in glv16 vdp harmonic glv256; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_r1_$w.json 2>>gpurun_out/bench_r1_final.err; done
python bench.py --reduce none --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r1_glv64_reduce_none.json 2>>gpurun_out/bench_r1_final.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],2), d.get('roofline',{}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"
tail -3 gpurun_out/bench_r1_final.err
