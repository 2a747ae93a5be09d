for sk in 0 100 250 500; do
  echo "skew ${sk}k clk"; SB200_QR_SKEW=$sk timeout 600 python bench.py --n 1048576 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['config']['solve_residual'])"
done
