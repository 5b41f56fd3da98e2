set -x
mkdir -p gpurun_out
for seg in 16 8 32; do VA_PAIR_SEG=$seg timeout 300 python bench.py --workload glv256 --ckpt-policy recompute --steps 3 --warmup 2 > gpurun_out/b256_seg$seg.json 2>&1; tail -c 200 gpurun_out/b256_seg$seg.json; done
timeout 300 python bench.py --workload glv256long --ckpt-policy recompute --steps 2 --warmup 1 > gpurun_out/b256long_seg.json 2>&1; tail -c 200 gpurun_out/b256long_seg.json
