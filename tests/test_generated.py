"""The PTX replay loops under evogp_b200/csrc/*.inc are generated files that are committed: they must be what
csrc/gen_fastpath.py produces today with its default settings (CPU test; nothing is compiled or launched)."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "evogp_b200", "csrc")
INCS = ["fastpath_k8.inc", "fastpath_k8_tmem.inc", "fastpath_k16_tmem.inc", "fastpath_k8_multi.inc"]


def test_committed_ptx_loops_are_what_the_generator_emits(tmp_path):
    shutil.copy(os.path.join(CSRC, "gen_fastpath.py"), tmp_path / "gen_fastpath.py")
    env = {k: v for k, v in os.environ.items() if not k.startswith("EVOGP_GEN_")}
    subprocess.run([sys.executable, "gen_fastpath.py"], cwd=tmp_path, env=env, check=True, capture_output=True)
    for name in INCS:
        fresh = (tmp_path / name).read_text()
        committed = open(os.path.join(CSRC, name)).read()
        assert fresh == committed, f"{name} is stale: run python evogp_b200/csrc/gen_fastpath.py"


def test_16_wide_loop_keeps_the_measured_layout():
    """The shipped 16-datapoint loop: a body per (form, operator) pair, pow / sinh / cosh through the generic interpreter,
    loss and pass loop outside the PTX block (profiles/r2_placement.md); the 8-datapoint loops: one body per rare operator."""
    k16 = open(os.path.join(CSRC, "fastpath_k16_tmem.inc")).read()
    k8 = open(os.path.join(CSRC, "fastpath_k8_tmem.inc")).read()
    assert "#define EVOGP_FASTPATH_K16_TMEM_ASM_FUSED_LOSS 0" in k16
    assert "L_CB:" not in k16 and "L_AV_MAX:" in k16 and "L_AV_POW" not in k16
    assert "L_CB:" in k8 and "L_B_POW:" in k8 and "L_U_SINH:" in k8 and "L_AV_MAX:" not in k8
    assert k8.count("brx.idx") == 3 and k16.count("brx.idx") == 1
