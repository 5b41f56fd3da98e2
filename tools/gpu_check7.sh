# occupancy experiment for the N = 16 kernel (build knob GLV_MINB16) + refreshed bench lines of the side workloads
set -x
mkdir -p gpurun_out
for v in 3 4; do VA_GLV_CTAS_PER_SM=$v timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 > gpurun_out/b16_c$v.json 2>&1; tail -c 250 gpurun_out/b16_c$v.json; done
timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 > gpurun_out/b16_m5.json 2>&1; tail -c 250 gpurun_out/b16_m5.json
for m in 6 8; do VA_ENGINE_LIB=$PWD/vectorizedadjoint_b200/libva_engine_m$m.so timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 > gpurun_out/b16_m$m.json 2>&1; tail -c 250 gpurun_out/b16_m$m.json; done
timeout 300 python bench.py --workload vdp --steps 3 --warmup 3 > gpurun_out/b_vdp.json 2>&1; tail -c 300 gpurun_out/b_vdp.json
timeout 300 python bench.py --workload harmonic --steps 3 --warmup 3 > gpurun_out/b_ho.json 2>&1; tail -c 300 gpurun_out/b_ho.json
timeout 300 python bench.py --workload glv256 --species 128 --steps 3 --warmup 2 > gpurun_out/b128_pair.json 2>&1; tail -c 250 gpurun_out/b128_pair.json
VA_GLV_NO_RING=1 timeout 300 python bench.py --workload glv256 --species 128 --steps 2 --warmup 1 > gpurun_out/b128_stream.json 2>&1; tail -c 250 gpurun_out/b128_stream.json
