for tool in memcheck racecheck; do echo "== $tool"; timeout 800 compute-sanitizer --tool $tool --print-limit 3 python tools/sanitize.py 2>&1 | grep -v "^$" | tail -9; done
