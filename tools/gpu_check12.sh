set -x
mkdir -p gpurun_out
timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 > gpurun_out/b16_w8.json 2>&1; tail -c 200 gpurun_out/b16_w8.json
for w in 4 6; do VA_ENGINE_LIB=$PWD/vectorizedadjoint_b200/libva_engine_w$w.so timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 > gpurun_out/b16_w$w.json 2>&1; tail -c 200 gpurun_out/b16_w$w.json; done
