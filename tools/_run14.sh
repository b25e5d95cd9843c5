mkdir -p gpurun_out/r2x
(timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2x/full_gpu_tests.log
for i in 1 2 3; do timeout 100 python tools/time_eval.py 2 8 2>&1 | tail -1 >> gpurun_out/r2x/time_eval.jsonl; done
FAST="+,-,*,/,loose_div,max,min,<,>,<=,>=,sin,cos,tan,tanh,log,loose_log,exp,inv,loose_inv,neg,abs,sqrt,loose_sqrt"
ALL="if,+,-,*,/,loose_div,pow,loose_pow,max,min,<,>,<=,>=,sin,cos,tan,sinh,cosh,tanh,log,loose_log,exp,inv,loose_inv,neg,abs,sqrt,loose_sqrt"
for spec in "6:+,-,*,/,sin,cos,tan" "5:$FAST" "4:$ALL"; do
  L=${spec%%:*}; fs=${spec#*:}
  TIME_FUNCS="$fs" TIME_LAYERS=$L timeout 100 python tools/time_eval.py 2 10 2>&1 | tail -1 >> gpurun_out/r2x/time_eval.jsonl
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x/bench_n1.json 2> gpurun_out/r2x/bench_n1.err
tail -c 600 gpurun_out/r2x/bench_n1.err
cat gpurun_out/r2x/full_gpu_tests.log
