# round 2: fence.proxy.async.global instead of fence.proxy.async in the headline kernel -- t8 vs t8s bit-identity over 20000 sets (t8s keeps the
# all-spaces fence), parity subset, A/B rate against libva_exp_fenceall.so
set -x
mkdir -p gpurun_out/r02_fence
timeout 150 python tools/t8s_check.py > gpurun_out/r02_fence/t8s_check.log 2>&1; tail -3 gpurun_out/r02_fence/t8s_check.log
bash tools/gpu_r2_ab.sh fence fenceall
