for P in 2 1; do
  rm -f vectorizedadjoint_b200/csrc/build/va_glv_wide.o
  make -s -C vectorizedadjoint_b200/csrc GLV_PAIR=$P -j8 > /dev/null
  echo "== PAIR=$P"
  timeout 600 python -m pytest tests -m gpu -q -k "glv" 2>&1 | tail -2
  python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
