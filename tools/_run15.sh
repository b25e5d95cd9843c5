mkdir -p gpurun_out/r2x
(timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2x/full_gpu_tests2.log
cat gpurun_out/r2x/full_gpu_tests2.log
