"""The five BASELINE.json configurations as model objects on synthetic catalogues.

One place for the parity tests (``tests/test_gpu_configs.py``), ``bench.py``'s ``configs`` block and
``tools/config_sweep.py``, so that the three measure the same thing.  Every builder returns
``(label, model, truth, n_walkers)``; ``truth`` feeds :func:`mcmc_dynamics_b200.synthetic.initial_ball`.

C1  ``example/data/test.csv`` analogue (``bin/run.py:146-259,487-491``): ConstantFit, ``v_sys`` fixed at 0,
    16 walkers.  The catalogue is the committed fixture ``tests/golden/c1_example_catalogue.npz``.
C2  10^4 stars, ModelFit, fixed centre, 128 walkers (``bin/run_tests.py:131-152`` with the centre fixed).
C3  10^5 stars, 30 % field contaminants, ``pmember`` column, ``SingleStars`` background over M = 2000
    Besancon-style velocities (``analysis/runner.py:96-103,272-286``), 256 walkers.
C3b the same catalogue with ``ModelFitGB`` (fitted Gaussian background, ``analysis/model.py:391-456``).
C4  3 x 10^5 stars, omega Cen-like: free centre, ``v_sys`` fixed at 232.5 km/s, bounds of
    ``bin/run_test_5139_center.py:157-165``, 128 walkers.
C5  10^7 stars x 1024 walkers, ModelFit (fixed or free centre): the headline sweep.
"""
import os

import numpy as np

from . import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C1_FIXTURE = os.path.join(ROOT, 'tests', 'golden', 'c1_example_catalogue.npz')

#: omega Cen centre used by bin/run_test_5139_center.py:48
OMEGA_CEN = (201.696718746, -47.479909445555)


def fix_centre(model, truth, free=False):
    model.parameters['ra_center'].set(value=truth['ra_center'], fixed=not free)
    model.parameters['dec_center'].set(value=truth['dec_center'], fixed=not free)


def config_c1(device=0):
    from .analysis import ConstantFit
    d = np.load(C1_FIXTURE)
    data = synthetic.reader_from_columns({k: d[k] for k in ('ra', 'dec', 'v', 'verr')})
    truth = {'ra_center': float(d['ra_center']), 'dec_center': float(d['dec_center']), 'v_sys': 0.0,
             'sigma_max': 30.0, 'v_maxx': 2.0, 'v_maxy': -2.0}
    m = ConstantFit(data, device=device)
    fix_centre(m, truth)
    m.parameters['v_sys'].set(value=0.0, fixed=True)        # bin/run.py:487-491
    return 'C1 example catalogue (6284 stars), ConstantFit, v_sys fixed', m, truth, 16


def config_c2(device=0):
    from .analysis import ModelFit
    data, truth = synthetic.mock_cluster(10_000, seed=1)
    m = ModelFit(data, device=device)
    fix_centre(m, truth)
    return 'C2 1e4 stars, ModelFit fixed centre', m, truth, 128


def c3_columns(n_stars=100_000):
    cols, truth = synthetic.mock_cluster(n_stars, seed=2, as_reader=False)
    cols, sample_field = synthetic.add_background(cols, truth, seed=102)
    truth = dict(truth, v_back=5.0, sigma_back=55.0, f_back=0.3)
    return cols, truth, sample_field


def config_c3(device=0, n_stars=100_000, m_background=2000):
    from .analysis import ModelFit
    from .background import SingleStars
    cols, truth, sample_field = c3_columns(n_stars)
    bg = SingleStars(sample_field(m_background, seed=202), device=device)
    m = ModelFit(synthetic.reader_from_columns(cols), background=bg, device=device)
    fix_centre(m, truth)
    return 'C3 %.0e stars, ModelFit + SingleStars(M=%d) mixture' % (n_stars, m_background), m, truth, 256


def config_c3b(device=0, n_stars=100_000):
    from .analysis import ModelFitGB
    cols, truth, _ = c3_columns(n_stars)
    m = ModelFitGB(synthetic.reader_from_columns(cols), device=device)
    fix_centre(m, truth)
    return 'C3b %.0e stars, ModelFitGB (fitted Gaussian background)' % n_stars, m, truth, 256


def config_c4(device=0, n_stars=300_000):
    from .analysis import ModelFit
    data, truth = synthetic.mock_cluster(n_stars, seed=3, ra_center=OMEGA_CEN[0], dec_center=OMEGA_CEN[1], v_sys=232.5)
    m = ModelFit(data, device=device)
    fix_centre(m, truth, free=True)
    m.parameters['v_sys'].set(value=232.5, fixed=True)      # bin/run_test_5139_center.py:163
    m.parameters['sigma_max'].set(min=0, max=100)            # bin/run_test_5139_center.py:157-165
    m.parameters['a'].set(min=0, max=300)
    m.parameters['v_maxx'].set(min=-100, max=100)
    m.parameters['v_maxy'].set(min=-100, max=100)
    m.parameters['r_peak'].set(min=0, max=500)
    return 'C4 %.0e stars, ModelFit free centre, omega Cen-like bounds, v_sys fixed 232.5' % n_stars, m, truth, 128


def config_c5(device=0, free=False, n_stars=10_000_000):
    from .analysis import ModelFit
    data, truth = synthetic.mock_cluster(n_stars, seed=4)
    m = ModelFit(data, device=device)
    fix_centre(m, truth, free=free)
    return 'C5 %.0e stars, ModelFit %s centre' % (n_stars, 'free' if free else 'fixed'), m, truth, 1024


BUILDERS = {'C1': config_c1, 'C2': config_c2, 'C3': config_c3, 'C3b': config_c3b, 'C4': config_c4, 'C5': config_c5}
