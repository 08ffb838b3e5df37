import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import numpy as np
import config_sweep as cs
from mcmc_dynamics_b200 import sampler as samplers, synthetic
name, model, truth, n_walkers = cs.config_c1()
theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
h = samplers.HostEnsembleSampler(n_walkers, model.n_fitted_parameters, model.lnprob, seed=1)
pos, lnp, _ = h.run_mcmc(theta, 50, store=False)
pr = cProfile.Profile(); pr.enable()
h.run_mcmc(pos, 2000, log_prob0=lnp, store=False)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:4500])
