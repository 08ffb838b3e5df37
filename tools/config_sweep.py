#!/usr/bin/env python
"""All five BASELINE.json configurations on one B200: parity against the NumPy oracle, terms/s of the
lnprob call (device-resident theta and through the host-buffer C ABI), emcee steps/s with the host
stretch-move loop and with the device-resident sampler, and the CPU oracle beside them.

    python tools/config_sweep.py [--out profiles/rNN_configs.md] [--quick]

This is a reporting tool (bench.py is the contract benchmark); it prints a Markdown table.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mcmc_dynamics_b200 import configs  # noqa: E402
from mcmc_dynamics_b200 import sampler as samplers  # noqa: E402
from mcmc_dynamics_b200 import synthetic  # noqa: E402
from mcmc_dynamics_b200.analysis import ConstantFit  # noqa: E402
from oracle import harness  # noqa: E402


def config_c3(gb=False):
    if gb:
        return configs.config_c3b()
    t0 = time.perf_counter()
    name, m, truth, w = configs.config_c3()
    return name + ' [construction incl. bg precompute %.0f ms]' % (1e3 * (time.perf_counter() - t0)), m, truth, w


config_c1, config_c2, config_c4, config_c5, fix_centre = (configs.config_c1, configs.config_c2, configs.config_c4,
                                                          configs.config_c5, configs.fix_centre)


def time_device(model, theta_dev, reps):
    for _ in range(3):
        model.lnprob_tensor(theta_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model.lnprob_tensor(theta_dev)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def time_host(model, theta, reps):
    for _ in range(3):
        model.lnprob(theta)
    t0 = time.perf_counter()
    for _ in range(reps):
        model.lnprob(theta)
    return (time.perf_counter() - t0) / reps


def run(name, model, truth, n_walkers, quick):
    n = model.n_data
    half = n_walkers // 2
    theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
    got = model.lnprob(theta[:half])
    n_check = min(half, 4 if n > 1_000_000 else 16)
    oracle = harness.oracle_for(model)
    t0 = time.perf_counter()
    want = oracle.lnprob_many(theta[:n_check])
    cpu_s = (time.perf_counter() - t0) / n_check
    err = harness.relative_error(got[:n_check], want)
    reps = 5 if n >= 1_000_000 else 50
    dev_s = time_device(model, torch.as_tensor(theta[:half], device='cuda:0'), reps)
    host_s = time_host(model, theta[:half], reps)
    terms = half * n
    # sampler steps/s (one step = two half-ensemble calls)
    steps = 20 if n >= 1_000_000 else (100 if quick else 300)
    s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=1)
    s.run_mcmc(theta, 5, store=False)
    t0 = time.perf_counter()
    s.run_mcmc(None, steps, store=False)
    dev_steps = steps / (time.perf_counter() - t0)
    h = samplers.HostEnsembleSampler(n_walkers, model.n_fitted_parameters, model.lnprob, seed=1)
    pos, lnp, _ = h.run_mcmc(theta, 3, store=False)
    t0 = time.perf_counter()
    h.run_mcmc(pos, steps, log_prob0=lnp, store=False)
    host_steps = steps / (time.perf_counter() - t0)
    cpu_terms = n / cpu_s
    return ('| %s | %d | %d | %.1e | %.3g | %.3g | %.3g | %.1f | %.1f | %.3g | %.2f |' % (
        name, n, n_walkers, err, terms / dev_s, terms / host_s, 1e6 * host_s, dev_steps, host_steps, cpu_terms,
        n_walkers * n / cpu_terms))


def bins_workflow(n_stars=3000, n_walkers=100, n_steps=100):
    """The per-radial-bin workflow of bin/run_tests.py:81-97 (100 walkers x 100 steps per bin):
    all bins in one segmented launch vs one device-sampler run per bin vs the host stretch move."""
    from mcmc_dynamics_b200.analysis import RadialBinsFit
    data, truth = synthetic.mock_cluster(n_stars, seed=6)
    data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=50, dlogr=0.1)
    fit = RadialBinsFit(data, model_class=ConstantFit)
    fix_centre(fit, truth)
    for name, expr in (('sigma_max', 'rng.lognormal(mean=2.3, sigma=0.5, size=n)'),
                       ('v_maxx', 'rng.normal(loc=0, scale=3, size=n)'), ('v_maxy', 'rng.normal(loc=0, scale=3, size=n)')):
        fit.parameters[name].set(initials=expr)
    pos = fit.get_initials(n_walkers)
    fit(n_walkers=n_walkers, n_steps=5, pos=pos, seed=1)                 # warm-up: pack, kernels loaded
    batched = 1e30
    for _ in range(3):                                                   # best of three (host-side jitter)
        t0 = time.perf_counter()
        fit(n_walkers=n_walkers, n_steps=n_steps, pos=pos, seed=1)
        batched = min(batched, time.perf_counter() - t0)
    models = [fit.bin_model(b) for b in range(fit.n_bins)]
    for m in models:
        m.pack()
    t0 = time.perf_counter()
    for b, m in enumerate(models):
        m(n_walkers=n_walkers, n_steps=n_steps, pos=pos[b], sampler='device', seed=1, prefix=None)
    per_bin_device = time.perf_counter() - t0
    t0 = time.perf_counter()
    for b, m in enumerate(models[:4]):
        m(n_walkers=n_walkers, n_steps=n_steps, pos=pos[b], sampler='host', seed=1, prefix=None)
    per_bin_host = (time.perf_counter() - t0) * fit.n_bins / 4.0
    oracle = harness.oracle_for(models[0])
    t0 = time.perf_counter()
    oracle.lnprob_many(pos[0])
    cpu = (time.perf_counter() - t0) * n_steps * fit.n_bins          # W lnprob calls per step, per bin
    return ('radial bins: %d stars in %d bins, %d walkers x %d steps per bin | batched device %.3f s | '
            'one device sampler per bin %.3f s | host stretch move per bin %.3f s | CPU oracle (1 process, '
            'extrapolated) %.1f s' % (n_stars, fit.n_bins, n_walkers, n_steps, batched, per_bin_device, per_bin_host, cpu))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--quick', action='store_true')
    args = ap.parse_args()
    lines = [
        '| config | stars | walkers | max rel. err vs oracle | terms/s (device theta) | terms/s (host buffers, C ABI) | '
        'µs per half-ensemble call (host) | steps/s device sampler | steps/s host sampler | CPU oracle terms/s '
        '(1 process) | CPU s per emcee step (1 process) |',
        '|---|---|---|---|---|---|---|---|---|---|---|']
    builders = [config_c1, config_c2, config_c3, lambda: config_c3(gb=True), config_c4, config_c5,
                lambda: config_c5(free=True)]
    if args.quick:
        builders = builders[:5]
    for build in builders:
        name, model, truth, n_walkers = build()
        line = run(name, model, truth, n_walkers, args.quick)
        print(line, flush=True)
        lines.append(line)
        model.pack().close()
        del model
    line = bins_workflow()
    print(line, flush=True)
    lines += ['', line]
    text = '\n'.join(lines) + '\n'
    if args.out:
        with open(args.out, 'w') as f:
            f.write('# BASELINE.json configurations on one B200 (tools/config_sweep.py)\n\n' + text)


if __name__ == '__main__':
    main()
