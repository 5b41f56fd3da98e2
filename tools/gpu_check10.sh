set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glv_t8 -c 1 -o gpurun_out/t8_small python bench.py --batch 16384 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_t8.log 2>&1; tail -3 gpurun_out/ncu_t8.log
