S=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; echo "bench wall $(( $(date +%s) - S )) s"; tail -c 200 gpurun_out/bench_r1_n1.json
