# developer probe: evaluation time by function set (tools/time_eval.py), 16 against 8 datapoints per lane
mkdir -p gpurun_out/r2w
ALL="if,+,-,*,/,loose_div,pow,loose_pow,max,min,<,>,<=,>=,sin,cos,tan,sinh,cosh,tanh,log,loose_log,exp,inv,loose_inv,neg,abs,sqrt,loose_sqrt"
FAST="+,-,*,/,loose_div,max,min,<,>,<=,>=,sin,cos,tan,tanh,log,loose_log,exp,inv,loose_inv,neg,abs,sqrt,loose_sqrt"
for k in 16 8; do
for fs in "$ALL" "$FAST" "+,-,*,/,sin,cos,tan,exp,log,sqrt,abs,neg" "pow,+,-,*,/"; do
  EVOGP_REPLAY_K=$k TIME_FUNCS="$fs" TIME_LAYERS=4 timeout 120 python tools/time_eval.py 2 20 2>&1 | tail -1 >> gpurun_out/r2w/funcs_k.jsonl
done
done
