"""In-tree build of the native pieces (no setuptools, no JIT cache):

  evogp_b200/lib/libevogp_b200.so   nvcc, sm_100a only — kernels + C ABI (include/evogp_b200.h)
  evogp_b200/lib/evogp_cuda_ops.so  g++ against torch — TORCH_LIBRARY(evogp_cuda) over the C ABI

Both are git-ignored and travel to the GPU box with the tree.  ``python -m evogp_b200.build``
rebuilds what is stale; ``--force`` rebuilds everything.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB_SO = os.path.join(LIBDIR, "libevogp_b200.so")
OPS_SO = os.path.join(LIBDIR, "evogp_cuda_ops.so")

CU_SOURCES = ["runtime.cu", "eval.cu", "eval_exchange.cu", "eval_acc.cu", "splice.cu", "generate.cu", "nextgen.cu", "select.cu", "host_api.cu"]


def _headers():
    """Everything a .cu may include: every .cuh / generated .inc in csrc/ (and the generator itself), plus the C ABI header."""
    import glob

    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc"))
                  + glob.glob(os.path.join(CSRC, "gen_*.py"))) + [os.path.normpath(os.path.join(CSRC, "../../include/evogp_b200.h"))]


# -use_fast_math: the reference's numeric contract (its setup.py passes the same flag), see DESIGN.md
NVCC_FLAGS = ["-O3", "-std=c++17", "-use_fast_math", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xptxas", "-O3"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    hdrs = _headers()
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(LIBDIR, os.path.basename(s).replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(LIB_SO, objs):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_SO] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB_SO


def build_torch_ops(force=False):
    import torch
    from torch.utils import cpp_extension as ce

    src = os.path.join(CSRC, "torch_ops.cpp")
    hdr = os.path.normpath(os.path.join(CSRC, "../../include/evogp_b200.h"))
    if not (force or _stale(OPS_SO, [src, hdr, LIB_SO])):
        return OPS_SO
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = []
    for p in ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True):
        inc += ["-isystem", p]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
           + inc + [src, "-o", OPS_SO, f"-L{LIBDIR}", "-levogp_b200", f"-L{torch_lib}", "-lc10", "-lc10_cuda",
                    "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"])
    subprocess.check_call(cmd)
    return OPS_SO


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_torch_ops(force)
    return LIB_SO, OPS_SO


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
