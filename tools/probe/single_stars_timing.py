"""SingleStars background precompute (background/single_stars.py:42-77): N x M pair evaluations per second."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from mcmc_dynamics_b200.background import SingleStars
from oracle import reference_np as ref
rng = np.random.default_rng(0)
for n, m in ((100_000, 2000), (10_000, 2000), (500, 300), (1_000_000, 2000)):
    v = rng.normal(0, 60, n); verr = 0.5 + 5 * rng.random(n); vbg = rng.normal(5, 55, m)
    bg = SingleStars(vbg)
    out = bg(v, verr)
    best = 1e30
    for _ in range(5):
        t0 = time.perf_counter(); out = bg(v, verr); best = min(best, time.perf_counter() - t0)
    k = min(n, 2000)
    want = ref.single_stars_background(vbg, v[:k], verr[:k])
    err = np.max(np.abs(out[:k] - want) / np.maximum(1, np.abs(want)))
    print('SingleStars N=%d M=%d: %.3f ms per call (host buffers) = %.3g pairs/s, max rel err vs oracle %.2e' % (
        n, m, 1e3 * best, n * m / best, err), flush=True)
