for cell in 0 0.08 0.16 0.25 0.4; do
  if [ $cell = 0 ]; then export OQG_SITE_GROUPING=0; else export OQG_SITE_GROUPING=1 OQG_SITE_CELL=$cell; fi
  python bench.py --ruptures 30000 --steps 2 --warmup 3 --no-cpu --no-strong 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cell $cell  %.4g evals/s  kernels ms %s  executed %.3g' % (d['value'], [round(x,2) for x in d['roofline']['kernel_ms']], (d['roofline'].get('executed') or {}).get('evals_per_step', 0)))"
done
