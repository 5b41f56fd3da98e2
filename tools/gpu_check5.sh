set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "glv256" > gpurun_out/t_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pair.log
tail -4 gpurun_out/t_pair.log
for cl in 2 4; do
VA_GLV_CLUSTER=$cl timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_cl$cl.json 2> gpurun_out/b256_cl$cl.err; tail -c 300 gpurun_out/b256_cl$cl.json
done
