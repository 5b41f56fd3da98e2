# round 2 profiling call: ncu --set full of one launch of a side workload's kernel (quad / scalar / pair / t8), after one warm-up launch
# usage: bash tools/gpu_r2_prof.sh <workload> <batch> <kernel regex> <tag>
set -x
W=$1; B=$2; K=$3; TAG=$4
mkdir -p gpurun_out/r02_prof
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_prof/$TAG \
  python bench.py --workload $W --batch $B --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-side --no-parity-sample > gpurun_out/r02_prof/$TAG.log 2>&1
tail -3 gpurun_out/r02_prof/$TAG.log
ls -la gpurun_out/r02_prof/
