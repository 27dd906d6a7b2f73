for f in variants/*.so; do
  echo "== $f"
  TSC_LIB=$PWD/$f python bench.py --mode sim --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms_per_launch'], d['value'])"
done
