set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glv_quad -c 1 -o gpurun_out/quad_full python bench.py --workload glv16 --batch 37888 --steps 1 --warmup 1 > gpurun_out/ncu_quad.log 2>&1; tail -3 gpurun_out/ncu_quad.log
