timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_t8.json 2>gpurun_out/bench_r1_t8.err; cut -c1-250 gpurun_out/bench_r1_t8.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_glv_t8 -s 2 -c 1 -f -o gpurun_out/prof_t8_c python bench.py --batch 16384 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-100
