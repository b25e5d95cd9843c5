mkdir -p gpurun_out/r2y
nvidia-smi --query-gpu=serial --format=csv,noheader >> gpurun_out/r2y/lottery_fused3.jsonl
for rep in 1 2 3 4 5 6; do
  timeout 100 python tools/time_eval.py 2 8 2>&1 | tail -1 >> gpurun_out/r2y/lottery_fused3.jsonl
  EVOGP_B200_LIB=build_variants/v_nofuse/libevogp_b200.so timeout 100 python tools/time_eval.py 2 8 2>&1 | tail -1 >> gpurun_out/r2y/lottery_fused3.jsonl
done
