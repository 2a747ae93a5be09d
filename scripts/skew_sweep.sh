for sk in 0 30000 60000 120000; do
  echo "skew $sk"; SB200_QR_SKEW=$sk timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['solve_residual'])"
done
