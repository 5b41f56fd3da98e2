timeout 600 python -m pytest tests -m gpu -x -q -k glv 2>&1 | tail -3
timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
