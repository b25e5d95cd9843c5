import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G
from evogp_b200 import _native
L = 8
# tree: ADD( DIV(1, 0), x0 )  prefix: [+, /, 1, 0, x0]
t = np.zeros((2, L), np.int16); v = np.zeros((2, L), np.float32); s = np.zeros((2, L), np.int16)
t[0, :5] = [3, 3, 1, 1, 0]; v[0, :5] = [1, 4, 1.0, 0.0, 0]; s[0, :5] = [5, 3, 1, 1, 1]
# tree: MUL(x0, SUB(1, 1))
t[1, :5] = [3, 0, 3, 1, 1]; v[1, :5] = [3, 0, 2, 1.0, 1.0]; s[1, :5] = [5, 1, 3, 1, 1]
dv, dt, ds = G.to_dev(v, t, s)
X = torch.tensor([[0.5], [2.0]], device="cuda"); y = torch.zeros(2, 1, device="cuda")
for fast in (0, 1):
    ws, n = G._ws(_native, 2, L)
    prog = torch.zeros((2, (L + 2) & ~1), dtype=torch.int64, device="cuda")
    rc = _native.abi().evogp_debug_lower(2, L, 1, 1, G._p(dv), G._p(dt), G._p(ds), fast, 0, G._p(ws), C.c_size_t(n), G._p(prog), G._stream())
    torch.cuda.synchronize()
    p = prog.cpu().numpy().view(np.uint64)
    for r in range(2):
        print("fast" if fast else "generic", r, [(hex(int(w & 0xFFFFFFFF)), hex(int(w >> 32))) for w in p[r][:5]])
print("fitness", G.abi_sr_fitness(_native, dv, dt, ds, X, y).cpu().numpy(), "fold env", os.environ.get("EVOGP_FOLD"))
