import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc
import gpu_util as G
from conftest import ALL_FUNCS, make_forest
from evogp_b200 import _native
_native.load_ops()
from evogp_b200.problem import Classification
from evogp_b200.tree import Forest
P, V, O, N, L = 1500, 6, 10, 300, 64
v, t, s = make_forest(orc, P, L, V, O, ALL_FUNCS, 4, keys=(61, 62), consts=(-1.0, 0.5, 2.0), out_prob=0.6)
rng = np.random.default_rng(3)
X = rng.normal(size=(N, V)).astype(np.float32); labels = rng.integers(0, 10, N).astype(np.float32)
dv, dt, ds, dX, dl = G.to_dev(v, t, s, X, labels)
prob = Classification(datapoints=dX, labels=dl, multi_output=True)
f = Forest(V, O, dv, dt, ds)
got = prob.evaluate(f); own = prob.evaluate_unfused(f)
out = f.batch_forward(dX)
p = torch.clip(torch.softmax(out, dim=2), 1e-15, 1 - 1e-15); pred_t = torch.argmax(p, dim=2)
nan = torch.isnan(out).any(2); mx = out.max(2).values
best = torch.where(torch.isnan(out), torch.full_like(out, -float("inf")), out)
pred_r = torch.argmax(best, 2)
pred_r = torch.where(nan | torch.isinf(mx), torch.zeros_like(pred_r), pred_r)
print("rule vs torch mismatch frac", (pred_r != pred_t).float().mean().item())
acc_r = (pred_r == dl[None, :]).float().mean(1)
print("got vs rule-on-outputs differ", (got != acc_r).float().mean().item(), "got vs own", (got != own).float().mean().item(), "own vs rule", (own != acc_r).float().mean().item())
bad = (pred_r != pred_t).nonzero()
for b in bad[:5]:
    i, n = int(b[0]), int(b[1])
    print(i, n, out[i, n].tolist(), p[i, n].tolist(), int(pred_t[i, n]), int(pred_r[i, n]))
d = (got != acc_r).nonzero()[:3]
for i in d[:, 0].tolist():
    print("tree", i, float(got[i]), float(acc_r[i]), float(own[i]))
