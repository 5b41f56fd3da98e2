set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "glv256 or large_species or checkpoint_policy" > gpurun_out/t_ring.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ring.log
tail -15 gpurun_out/t_ring.log
for f in 2 0 6 4; do
  VA_RING_FLAGS=$f timeout 300 python bench.py --workload glv256 --steps 3 --warmup 2 > gpurun_out/b256_f$f.json 2> gpurun_out/b256_f$f.err; tail -c 1500 gpurun_out/b256_f$f.json
done
VA_RING_FLAGS=2 timeout 300 python bench.py --workload glv256 --steps 2 --warmup 1 --reduce none > gpurun_out/b256_none.json 2>&1; tail -c 600 gpurun_out/b256_none.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_glv_ring -c 1 -o gpurun_out/ring_full python bench.py --workload glv256 --batch 1184 --steps 1 --warmup 1 > gpurun_out/ncu_ring.log 2>&1; tail -3 gpurun_out/ncu_ring.log
