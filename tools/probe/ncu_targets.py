"""Small drivers for the ncu captures of round 2 (one kernel each, a handful of launches):
    python tools/probe/ncu_targets.py mix | mixgb | c3 | c4 | single_stars | c5s
"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
what = sys.argv[1]
if what == 'single_stars':
    from mcmc_dynamics_b200.background import SingleStars
    rng = np.random.default_rng(0)
    n, m = 100_000, 2000
    v = rng.normal(0, 60, n); verr = 0.5 + 5 * rng.random(n); bg = SingleStars(rng.normal(5, 55, m))
    for _ in range(4):
        out = bg(v, verr)
    print('single_stars', float(out[0]))
else:
    sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
    import ab_configs
    from mcmc_dynamics_b200 import synthetic
    name, model, truth, nw = ab_configs.BUILDERS[what]()
    th = synthetic.initial_ball(truth, model.fitted_parameters, nw, seed=5, scale=0.05)[:nw // 2]
    tdev = torch.as_tensor(th, device='cuda:0')
    n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6       # a large count keeps a GPU busy as background load
    for k in range(n_calls):
        out = model.lnprob_tensor(tdev)
        if k % 64 == 63:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(name, float(out[0]))
