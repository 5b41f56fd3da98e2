python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2>> gpurun_out/bench_r1.err; cat gpurun_out/bench_ref_r1.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_glv_wide -s 1 -c 1 -o gpurun_out/prof_glv_r1_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
