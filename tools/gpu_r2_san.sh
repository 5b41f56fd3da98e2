# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over tools/sanitize.py: every kernel family, small runs
set -x
O=gpurun_out/r02_san2
mkdir -p $O
for tool in memcheck racecheck synccheck initcheck; do
  start=$(date +%s)
  timeout 700 compute-sanitizer --tool $tool python tools/sanitize.py > $O/san_$tool.log 2>&1; echo "$tool rc=$? $(( $(date +%s) - start )) s"; tail -2 $O/san_$tool.log
done
