# cluster kernel: parity subset + throughput under both checkpoint policies
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "recompute or auto_policy or glv256 or large_species or checkpoint_policy" > gpurun_out/t_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pair.log; tail -6 gpurun_out/t_pair.log
timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_pair.json 2>&1; tail -c 200 gpurun_out/b256_pair.json
timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 --reduce none > gpurun_out/b256_pair_none.json 2>&1; tail -c 200 gpurun_out/b256_pair_none.json
timeout 300 python bench.py --workload glv256 --ckpt-policy recompute --steps 3 --warmup 2 > gpurun_out/b256_seg16.json 2>&1; tail -c 200 gpurun_out/b256_seg16.json
VA_GLV_CLUSTER=4 timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_cl4.json 2>&1; tail -c 200 gpurun_out/b256_cl4.json
