timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -2
timeout 600 python tools/configs_report.py > gpurun_out/configs4.json 2> gpurun_out/configs4.err; tail -2 gpurun_out/configs4.err
