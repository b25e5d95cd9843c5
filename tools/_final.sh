timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lower_kernel|replay_kernel" -c 2 -s 6 -o gpurun_out/final3 python bench.py --steps 2 --warmup 2 --no-cpu --no-ref-gpu > gpurun_out/final3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final3.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ncu4.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -c 300 gpurun_out/bench_r1_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1_n1.err; tail -c 200 gpurun_out/bench_r1_ref.json
timeout 600 python tools/configs_report.py > gpurun_out/configs3.json 2> gpurun_out/configs3.err; tail -2 gpurun_out/configs3.err
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
ls -la gpurun_out/final3*
