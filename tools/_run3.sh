for v in 16 16; do
EVOGP_REPLAY_K=$v timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu --no-ref-gpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K=$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
