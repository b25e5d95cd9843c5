"""-m gpu, >= 2 GPUs: spawns tests/multi_gpu_check.py under torch.distributed.run (one process per GPU, NCCL) and
keeps its log.  Skipped on a single-GPU box; the driver-visible multi-GPU correctness check also runs inside bench.py
at --gpus N > 1 (`exchange_check`)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_sharded_fitness_and_replicated_populations(native, world):
    import torch

    have = torch.cuda.device_count()
    if have < world:
        pytest.skip(f"needs {world} GPUs, {have} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_n{world}.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-20000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-gpu check ok" in r.stdout
