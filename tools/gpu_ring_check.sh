# one GPU call: parity of the N = 256 kernels, their throughput, one ncu capture of the cluster kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "glv256 or large_species or checkpoint_policy" > gpurun_out/t_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pair.log
tail -15 gpurun_out/t_pair.log
for cl in 2 4; do
VA_DEBUG=1 VA_GLV_CLUSTER=$cl timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_cl$cl.json 2> gpurun_out/b256_cl$cl.err; tail -c 500 gpurun_out/b256_cl$cl.json; grep "va:" gpurun_out/b256_cl$cl.err | tail -2
VA_GLV_CLUSTER=$cl timeout 300 python bench.py --workload glv256 --steps 2 --warmup 1 --reduce none > gpurun_out/b256_cl${cl}_none.json 2>&1; tail -c 300 gpurun_out/b256_cl${cl}_none.json
done
VA_GLV_CLUSTER=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glv_pair -c 1 -o gpurun_out/pair4_full python bench.py --workload glv256 --batch 1184 --steps 1 --warmup 1 > gpurun_out/ncu_pair4.log 2>&1; tail -3 gpurun_out/ncu_pair4.log
