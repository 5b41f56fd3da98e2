set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "recompute or auto_policy or glv256 or large_species or checkpoint_policy" > gpurun_out/t_seg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_seg.log; tail -25 gpurun_out/t_seg.log
timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_store.json 2>&1; tail -c 250 gpurun_out/b256_store.json
