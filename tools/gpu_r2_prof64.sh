set -x
mkdir -p gpurun_out/r02_prof
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_glv_t8 --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_prof/$1 \
  python bench.py --traffic-probe > gpurun_out/r02_prof/$1.log 2>&1
tail -2 gpurun_out/r02_prof/$1.log
