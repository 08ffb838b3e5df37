"""Where a host-buffer lnprob call spends its time on the small configurations: the C entry point alone (ctypes
with pre-converted pointers) against the full Python path (Runner.lnprob -> PackedModel.lnprob -> ctypes)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from mcmc_dynamics_b200 import configs, synthetic
for key in ('C1', 'C2', 'C3'):
    name, model, truth, nw = configs.BUILDERS[key]()
    half = nw // 2
    theta = np.ascontiguousarray(synthetic.initial_ball(truth, model.fitted_parameters, nw, seed=5, scale=0.05)[:half])
    packed = model.pack()
    lib = packed._lib
    out = np.empty(half)
    tp, op, h = theta.ctypes.data, out.ctypes.data, packed.handle
    for _ in range(50): lib.mcd_lnprob(h, tp, half, op)
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n): lib.mcd_lnprob(h, tp, half, op)
    raw = (time.perf_counter() - t0) / n
    for _ in range(50): model.lnprob(theta)
    t0 = time.perf_counter()
    for _ in range(n): model.lnprob(theta)
    full = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n): packed.lnprob(theta)
    mid = (time.perf_counter() - t0) / n
    print('%s: C entry point alone %.1f us | PackedModel.lnprob %.1f us | Runner.lnprob %.1f us' % (key, 1e6 * raw, 1e6 * mid, 1e6 * full), flush=True)
