import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import numpy as np
from mcmc_dynamics_b200 import synthetic
from mcmc_dynamics_b200.analysis import ConstantFit, RadialBinsFit
import config_sweep as cs
data, truth = synthetic.mock_cluster(3000, seed=6)
data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=50, dlogr=0.1)
fit = RadialBinsFit(data, model_class=ConstantFit)
cs.fix_centre(fit, truth)
for name, expr in (('sigma_max', 'rng.lognormal(mean=2.3, sigma=0.5, size=n)'),
                   ('v_maxx', 'rng.normal(loc=0, scale=3, size=n)'), ('v_maxy', 'rng.normal(loc=0, scale=3, size=n)')):
    fit.parameters[name].set(initials=expr)
pos = fit.get_initials(100)
fit(n_walkers=100, n_steps=5, pos=pos, seed=1)
for rep in range(8):
    t0 = time.perf_counter()
    engine = fit(n_walkers=100, n_steps=100, pos=pos, seed=1)
    dt = time.perf_counter() - t0
    print('rep %d: %.2f ms, engine %s, bins %d' % (rep, 1e3 * dt, getattr(engine, 'engine', None), fit.n_bins))
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
fit(n_walkers=100, n_steps=100, pos=pos, seed=1)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(14); print(s.getvalue()[:3000])
