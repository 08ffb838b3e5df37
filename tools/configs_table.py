#!/usr/bin/env python
"""Markdown table of the `configs` block (BASELINE.json configurations C1-C4) and the headline of a bench.py line:

    python tools/configs_table.py profiles/r02_bench_n1.json > profiles/r02_configs.md
"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print('# BASELINE.json configurations on one B200 — `bench.py` (`configs` block of `%s`)\n' % sys.argv[1])
print('Each row is measured in-process by the driver-run benchmark: parity against the NumPy oracle on the first walkers '
      'of the half-ensemble, device time of the half-ensemble `mcd_lnprob_device` call (CUDA events, theta resident), '
      'the same call through the host-buffer C ABI (`Runner.lnprob`), measured emcee steps/s of the device-resident '
      'and the host stretch-move samplers, and the single-process NumPy oracle (the reference\'s arithmetic) beside them.\n')
print('| config | stars | walkers | max rel. err vs oracle | terms/s (device theta) | µs per call | terms/s (host buffers) | '
      'µs per call | steps/s device sampler (engine) | steps/s host sampler | steps/s CPU oracle | CPU terms/s |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|')
for key, c in d.get('configs', {}).items():
    s = c['steps_per_s']
    print('| %s: %s | %d | %d | %.1e | %.3g | %.1f | %.3g | %.1f | %.0f (%s) | %.0f | %.3g | %.3g |' % (
        key, c['workload'], c['n_stars'], c['n_walkers'], c['max_rel_err_vs_oracle'], c['terms_per_s'], c['us_per_call'],
        c['e2e_terms_per_s'], c['e2e_us_per_call'], s['device'], s['device_engine'], s['host'], s['cpu'],
        c['cpu_terms_per_s']))
s = d['steps_per_s']
print('| C5: %s | %d | %d | %.1e (2e5-star prefix) | %.3g | %.0f | %.3g | %.0f | %.1f (%s; steady state %.1f) | %.1f | %.3g | %.3g |' % (
    d['config']['workload'], d['config']['n_stars'], d['config']['n_walkers'], d['checks']['vs_oracle'], d['value'],
    1e3 * d['ms_per_step'] / 2, d['e2e']['value'], 1e3 * d['e2e']['ms_per_step'] / 2, s.get('device', float('nan')),
    s.get('device_engine', 'graph'), s.get('device_steady_state', float('nan')), s.get('host', float('nan')),
    s.get('cpu', float('nan')), d.get('cpu_baseline', {}).get('value', float('nan'))))
print('\nC5 CPU figures: %s' % d.get('cpu_baseline', {}).get('sample', 'n/a'))
