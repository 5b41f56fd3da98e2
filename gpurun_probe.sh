timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1b_final.json 2>gpurun_out/bench_r1b_final.err; cut -c1-300 gpurun_out/bench_r1b_final.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glv_t8 -s 1 -c 1 -f -o gpurun_out/prof_t8_full1M python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1; tail -3 gpurun_out/launches_r1b.csv
