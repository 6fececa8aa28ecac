for v in "$@"; do
  echo "== variant $v"
  IGLB200_LIBRARY=$PWD/libigl_b200/_variants/lib_$v.so python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['kernel_ms_per_step'])"
done
