"""Device sampler (graph engine) on the headline workload: steps/s with and without chain storage, direct and
through Runner.__call__."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from mcmc_dynamics_b200 import configs, synthetic
from mcmc_dynamics_b200 import sampler as samplers
n = int(os.environ.get('N_STARS', 10_000_000))
name, model, truth, nw = configs.config_c5(n_stars=n)
theta = synthetic.initial_ball(truth, model.fitted_parameters, nw, seed=5)
packed = model.pack()
half = theta[:nw // 2]
for _ in range(3): model.lnprob(half)
t0 = time.perf_counter()
for _ in range(10): model.lnprob(half)
print('lnprob half-ensemble call: %.3f ms' % (1e3 * (time.perf_counter() - t0) / 10), flush=True)
s = samplers.DeviceEnsembleSampler(nw, model.n_fitted_parameters, packed, seed=1)
t0 = time.perf_counter(); s.run_mcmc(theta, 5, store=False); print('first 5 steps (set_state, capture): %.1f ms' % (1e3 * (time.perf_counter() - t0)), flush=True)
for store in (False, True, False):
    for steps in (10, 30):
        t0 = time.perf_counter(); s.run_mcmc(None, steps, store=store); dt = time.perf_counter() - t0
        print('store=%s steps=%d: %.2f ms/step (%.1f steps/s), engine %s, acceptance %.3f' % (store, steps, 1e3 * dt / steps, steps / dt, s.engine, s.acceptance_fraction.mean()), flush=True)
t0 = time.perf_counter(); run = model(n_walkers=nw, n_steps=30, pos=theta, sampler='device', seed=1, prefix=None); dt = time.perf_counter() - t0
print('Runner.__call__ device 30 steps: %.2f ms/step' % (1e3 * dt / 30), flush=True)
t0 = time.perf_counter(); run = model(n_walkers=nw, n_steps=30, pos=theta, sampler='host', seed=1, prefix=None); dt = time.perf_counter() - t0
print('Runner.__call__ host 30 steps: %.2f ms/step' % (1e3 * dt / 30), flush=True)
