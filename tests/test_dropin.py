"""-m gpu: the drop-in claim, exercised.  The reference's own python front-end (baseline/_ref/evogp, unmodified) runs a
seeded GP loop twice in fresh processes — once over its own CUDA extension, once over this repo's operator library
swapped in at evogp/tree/__init__.py:2 — and every generation must agree: populations bit for bit on the valid
prefixes (integer / index work), fitness within 1e-5 relative (BASELINE.json north_star)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
GENS = 12


@pytest.fixture(scope="module")
def runs(native, tmp_path_factory):
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "evogp")):
        pytest.skip("baseline/_ref not installed (pip install --target baseline/_ref /root/reference)")
    d = tmp_path_factory.mktemp("dropin")
    out = {}
    for mode in ("reference", "ours"):
        path = str(d / f"{mode}.npz")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_run.py"), mode, path, "600", str(GENS)],
                           cwd=str(d), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, f"{mode} run failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
        out[mode] = np.load(path)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dropin.log"), "w") as f:
        f.write("reference native: %s\nours native: %s\n" % (list(out["reference"]["native"]), list(out["ours"]["native"])))
    return out


def close(a, b, rtol=1e-5):
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    ia = np.isinf(a)
    if not (np.array_equal(ia, np.isinf(b)) and np.array_equal(a[ia], b[ia])):
        return False
    m = ~(na | ia)
    return bool((np.abs(a[m].astype(np.float64) - b[m]) <= rtol * np.abs(b[m].astype(np.float64))).all())


def test_the_two_runs_used_different_native_code(runs):
    assert any(str(n).startswith("evogp_cuda.cpython") for n in runs["reference"]["native"])
    assert "evogp_cuda_ops.so" in list(runs["ours"]["native"]) and "libevogp_b200.so" in list(runs["ours"]["native"])


def test_seeded_loop_is_identical_under_the_reference_front_end(runs):
    ref, ours = runs["reference"], runs["ours"]
    for g in range(GENS):
        for name in ("type", "size", "value"):
            a, b = ref[f"{name}{g}"], ours[f"{name}{g}"]
            same = np.array_equal(a.view(np.uint32), b.view(np.uint32)) if a.dtype.kind == "f" else np.array_equal(a, b)
            assert same, f"generation {g}: node_{name} differs in {(a != b).any(1).sum()} of {a.shape[0]} trees"
        assert close(ours[f"fitness{g}"], ref[f"fitness{g}"]), f"generation {g}: fitness beyond 1e-5 relative"
        # per-tree outputs (Forest.batch_forward -> tree_evaluate) carry the same bits under both libraries, so the torch-mode
        # fitness that drives selection in both runs is identical, NaNs included
        ta, tb = ref[f"torch_fitness{g}"], ours[f"torch_fitness{g}"]
        assert np.array_equal(np.isnan(ta), np.isnan(tb)) and np.array_equal(ta[~np.isnan(ta)], tb[~np.isnan(tb)]), \
            f"generation {g}: torch-mode fitness differs in {(ta != tb).sum()} trees"


def test_forward_paths_and_pareto_front_agree(runs):
    ref, ours = runs["reference"], runs["ours"]
    assert close(ours["best_forward"], ref["best_forward"])
    assert close(ours["forest_forward"], ref["forest_forward"])
    assert close(ours["pareto_fitness"], ref["pareto_fitness"])
