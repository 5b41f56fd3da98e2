for MB in 3 2; do
  rm -f vectorizedadjoint_b200/csrc/build/va_glv_wide.o
  make -s -C vectorizedadjoint_b200/csrc GLV_LG=8 GLV_MINB=$MB -j8 > /dev/null
  echo "== MINB=$MB"
  python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
