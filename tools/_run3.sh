timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -2
for v in 1 2; do
timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu --no-ref-gpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
