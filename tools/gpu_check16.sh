set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "recompute or auto_policy or checkpoint_policy" > gpurun_out/t_seg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_seg.log; tail -4 gpurun_out/t_seg.log
for seg in 16 32; do VA_PAIR_SEG=$seg timeout 300 python bench.py --workload glv256 --ckpt-policy recompute --steps 3 --warmup 2 > gpurun_out/b256_seg$seg.json 2>&1; tail -c 200 gpurun_out/b256_seg$seg.json; done
timeout 300 python bench.py --workload glv256long --ckpt-policy recompute --steps 2 --warmup 1 > gpurun_out/b256long_seg.json 2>&1; tail -c 200 gpurun_out/b256long_seg.json
VA_GLV_NO_RING=1 timeout 300 python bench.py --workload glv256 --ckpt-policy recompute --batch 2048 --steps 2 --warmup 1 > gpurun_out/b256_stream_rec.json 2>&1; tail -c 200 gpurun_out/b256_stream_rec.json
