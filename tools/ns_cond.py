import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from rlitg_b200 import pirl_env as P, synth
dev = torch.device('cuda:0')
for seed in range(6):
    B = 20000
    q = synth.random_configs(B, seed=100 + seed)
    rng = np.random.default_rng(seed)
    err = rng.normal(0, 0.3, size=(B, 7))
    s = np.linalg.svd(O.jacobian(q), compute_uv=False)
    cond = s[:, 0] / s[:, -1]
    want = O.nullspace_norm(err, q)
    en = np.linalg.norm(err, axis=1)
    for dt, eps in ((torch.float32, 6e-8), (torch.float64, 1.1e-16)):
        got = P.nullspace_norm(torch.tensor(err, dtype=dt, device=dev), torch.tensor(q, dtype=dt, device=dev)).double().cpu().numpy()
        d = np.abs(got - want)
        r1 = d / (eps * cond * en)
        r2 = d / (eps * cond ** 2 * en)
        i = np.argmax(d)
        print(seed, dt, "max err %.3g at cond %.3g | max err/(eps cond |e|) %.3g | max err/(eps cond^2 |e|) %.3g | cond max %.3g" % (d[i], cond[i], r1.max(), r2.max(), cond.max()))
