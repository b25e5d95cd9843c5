timeout 600 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 600 gpurun_out/bench_r1_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1_n1.err; tail -c 400 gpurun_out/bench_r1_ref.json
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
