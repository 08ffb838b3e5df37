import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mcmc_dynamics_b200 import synthetic
from mcmc_dynamics_b200.analysis import ModelFit, ModelFitGB
from mcmc_dynamics_b200.background import Gaussian
n = 2_000_000
cols, truth = synthetic.mock_cluster(n, seed=2, as_reader=False)
cols, _ = synthetic.add_background(cols, truth, seed=102)
truth = dict(truth, v_back=5.0, sigma_back=55.0, f_back=0.3)
for name, make in (('ModelFit + fixed background (pmember)', lambda d: ModelFit(d, background=Gaussian(5.0, 55.0))),
                   ('ModelFitGB', lambda d: ModelFitGB(d))):
    m = make(synthetic.reader_from_columns(cols))
    m.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    m.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    # host-buffer C ABI (ctypes): the path that honours MCD_B200_LIB for A/B runs of kernel builds
    th = synthetic.initial_ball(truth, m.fitted_parameters, 512, seed=5, scale=0.05)
    for _ in range(20): m.lnprob(th)
    ms = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(20): m.lnprob(th)
        ms = min(ms, (time.perf_counter() - t0) / 20 * 1e3)
    print('%-40s %.3f ms per 512-walker call over %d stars = %.3g terms/s' % (name, ms, n, 512 * n / ms * 1e3))
