set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "quad or glv16 or reference_example or synthetic_vs_reference or batch_vs_oracle or half_norm or device_buffers" > gpurun_out/t_quad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_quad.log
tail -30 gpurun_out/t_quad.log
timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 > gpurun_out/b16_quad.json 2> gpurun_out/b16_quad.err; tail -c 400 gpurun_out/b16_quad.json; tail -3 gpurun_out/b16_quad.err
timeout 300 python bench.py --workload glv16 --steps 3 --warmup 2 --reduce none > gpurun_out/b16_quad_none.json 2>&1; tail -c 300 gpurun_out/b16_quad_none.json
