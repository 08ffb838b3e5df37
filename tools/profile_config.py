#!/usr/bin/env python
"""Launch the lnprob kernel of one BASELINE configuration a few times (for ncu):
    python tools/profile_config.py c3 [--calls 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import torch  # noqa: E402

import config_sweep as cs  # noqa: E402
from mcmc_dynamics_b200 import synthetic  # noqa: E402

CONFIGS = {'c1': cs.config_c1, 'c2': cs.config_c2, 'c3': cs.config_c3, 'c3b': lambda: cs.config_c3(gb=True),
           'c4': cs.config_c4, 'c5': cs.config_c5, 'c5free': lambda: cs.config_c5(free=True)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config', choices=sorted(CONFIGS))
    ap.add_argument('--calls', type=int, default=10)
    args = ap.parse_args()
    name, model, truth, n_walkers = CONFIGS[args.config]()
    theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
    th = torch.as_tensor(theta[:n_walkers // 2], device='cuda:0')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        out = model.lnprob_tensor(th)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.calls):
        out = model.lnprob_tensor(th)
    e1.record()
    torch.cuda.synchronize()
    info = model.pack().info()
    print(name, '| %.1f us per call | grid %d x %d, walkers per CTA %d | lnprob[0] = %.6f' % (
        1e3 * e0.elapsed_time(e1) / args.calls, info['last_grid_x'], info['last_grid_y'], info['last_walker_tile'],
        float(out[0])))


if __name__ == '__main__':
    main()
