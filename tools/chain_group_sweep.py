#!/usr/bin/env python
"""Steps/s of the resident-chain sampler engine against the number of CTAs that share one segment's
stars (MCD_CHAIN_GROUP), for the small and medium BASELINE configurations: calibrates the cost model of
``launch_resident_chain`` (csrc/mcd_api.cu).

    python tools/chain_group_sweep.py [--configs c1,c2,c3,c4] [--groups 1,2,4,8,16,37,74,148]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import config_sweep  # noqa: E402
from mcmc_dynamics_b200 import sampler as samplers  # noqa: E402
from mcmc_dynamics_b200 import synthetic  # noqa: E402


def steps_per_second(model, theta, n_walkers, steps):
    s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=1)
    s.run_mcmc(theta, 5, store=False)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        s.run_mcmc(None, steps, store=False)
        best = max(best, steps / (time.perf_counter() - t0))
    engine = s.engine
    s.close()
    return best, engine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c1,c2,c3,c4')
    ap.add_argument('--groups', default='1,2,4,8,16,37,74,148')
    ap.add_argument('--steps', type=int, default=400)
    args = ap.parse_args()
    builders = {'c1': config_sweep.config_c1, 'c2': config_sweep.config_c2, 'c3': config_sweep.config_c3,
                'c3b': lambda: config_sweep.config_c3(gb=True), 'c4': config_sweep.config_c4}
    for key in args.configs.split(','):
        name, model, truth, n_walkers = builders[key]()
        theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
        print('## %s (%d stars, %d walkers)' % (name, model.n_data, n_walkers), flush=True)
        os.environ['MCD_FORCE_RESIDENT_CHAIN'] = '0'
        os.environ.pop('MCD_CHAIN_GROUP', None)
        os.environ['MCD_NO_RESIDENT_CHAIN'] = '1'
        rate, engine = steps_per_second(model, theta, n_walkers, args.steps)
        print('graph of launches: %.0f steps/s' % rate, flush=True)
        os.environ['MCD_NO_RESIDENT_CHAIN'] = '0'
        rate, engine = steps_per_second(model, theta, n_walkers, args.steps)
        print('library choice %s: %.0f steps/s' % (engine, rate), flush=True)
        os.environ['MCD_FORCE_RESIDENT_CHAIN'] = '1'
        for g in args.groups.split(','):
            os.environ['MCD_CHAIN_GROUP'] = g
            rate, engine = steps_per_second(model, theta, n_walkers, args.steps)
            print('group %s -> %s: %.0f steps/s (%.2f us per half-step)' % (g, engine, rate, 0.5e6 / rate), flush=True)
        os.environ.pop('MCD_CHAIN_GROUP', None)
        model.pack().close()


if __name__ == '__main__':
    main()
