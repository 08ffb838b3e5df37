"""Shared builders for the parity tests: one seeded catalogue per variant, product model + oracle twin."""
import numpy as np

from mcmc_dynamics_b200 import synthetic
from mcmc_dynamics_b200.analysis import (ConstantFit, ConstantFitGB, ModelFit, ModelFitGB,
                                         ModelFitConstantBackground)
from mcmc_dynamics_b200.background import Gaussian, SingleStars
from oracle import harness
from oracle import reference_np as ref

#: tolerance of north_star: lnprob within 1e-9 relative in FP64
RTOL = 1e-9

VARIANTS = ['ConstantFit', 'ConstantFit+bg', 'ConstantFitGB', 'ModelFit', 'ModelFit+bg', 'ModelFitGB',
            'ModelFitConstantBackground']


def build(variant, n_stars=3000, free_centre=False, seed=1, math_mode='fast', v_sys=0.0):
    """Returns (model, oracle, theta_sampler(n_walkers))."""
    columns, truth = synthetic.mock_cluster(n_stars, seed=seed, as_reader=False, v_sys=v_sys)
    background = None
    oracle_lbg = None
    if variant != 'ConstantFit' and variant != 'ModelFit':
        columns, sample_field = synthetic.add_background(columns, truth, seed=seed + 100)
        v_bg = sample_field(300, seed=seed + 200)
    data = synthetic.reader_from_columns(columns)
    kw = dict(math_mode=math_mode)
    if variant == 'ConstantFit':
        model = ConstantFit(data, **kw)
    elif variant == 'ModelFit':
        model = ModelFit(data, **kw)
    elif variant == 'ConstantFit+bg':
        model = ConstantFit(data, background=SingleStars(v_bg), **kw)
        oracle_lbg = ref.single_stars_background(v_bg, columns['v'], columns['verr'])
    elif variant == 'ModelFit+bg':
        model = ModelFit(data, background=Gaussian(5.0, 55.0), **kw)
        oracle_lbg = ref.gaussian_background(columns['v'], columns['verr'], 5.0, 55.0)
    elif variant == 'ConstantFitGB':
        model = ConstantFitGB(data, **kw)
    elif variant == 'ModelFitGB':
        model = ModelFitGB(data, **kw)
    elif variant == 'ModelFitConstantBackground':
        model = ModelFitConstantBackground(data, background=SingleStars(v_bg), **kw)
        oracle_lbg = ref.single_stars_background(v_bg, columns['v'], columns['verr'])
    else:
        raise ValueError(variant)
    truth = dict(truth, v_back=5.0, sigma_back=55.0, f_back=0.3)
    if not free_centre:
        model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
        model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    else:
        model.parameters['ra_center'].set(value=truth['ra_center'])
        model.parameters['dec_center'].set(value=truth['dec_center'])
    oracle = harness.oracle_for(model, lnlike_background=oracle_lbg)

    def theta(n_walkers, seed=7, scale=0.2):
        return synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=seed, scale=scale)

    return model, oracle, theta, truth
