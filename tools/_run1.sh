timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -5
for t in 0 1; do
EVOGP_TMEM_STACK=$t timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu --no-ref-gpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tmem=$t', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
