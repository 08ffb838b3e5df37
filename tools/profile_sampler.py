#!/usr/bin/env python
"""Run the device-resident ensemble sampler of one BASELINE configuration (for ncu launch lists):
    python tools/profile_sampler.py c2 [--steps 400]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import config_sweep as cs  # noqa: E402
from mcmc_dynamics_b200 import sampler as samplers  # noqa: E402
from mcmc_dynamics_b200 import synthetic  # noqa: E402

CONFIGS = {'c1': cs.config_c1, 'c2': cs.config_c2, 'c3': cs.config_c3, 'c3b': lambda: cs.config_c3(gb=True),
           'c4': cs.config_c4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config', choices=sorted(CONFIGS))
    ap.add_argument('--steps', type=int, default=400)
    args = ap.parse_args()
    name, model, truth, n_walkers = CONFIGS[args.config]()
    theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
    s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=1)
    s.run_mcmc(theta, 5, store=False)
    t0 = time.perf_counter()
    s.run_mcmc(None, args.steps, store=True)
    dt = time.perf_counter() - t0
    print('%s | %d steps with stored chain in %.2f ms = %.0f steps/s | engine %s | acceptance %.2f' % (
        name, args.steps, 1e3 * dt, args.steps / dt, s.engine, s.acceptance_fraction.mean()))


if __name__ == '__main__':
    main()
