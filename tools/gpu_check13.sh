set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 40 --csv --log-file gpurun_out/vdp_launches.csv python bench.py --workload vdp --steps 2 --warmup 1 > gpurun_out/vdp_ncu.log 2>&1; tail -2 gpurun_out/vdp_ncu.log
timeout 900 python -m pytest tests -x -q -m gpu -k "quad or glv16 or torch or abi" > gpurun_out/t_last.log 2>&1; tail -3 gpurun_out/t_last.log
