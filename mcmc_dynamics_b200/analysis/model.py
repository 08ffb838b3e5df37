"""Radial-profile fits (``mcmc_dynamics/analysis/model.py``).

``ModelFit`` (``model.py:20-335``): Lynden-Bell rotation curve
``v_los = v_sys + 2 (v_max / r_peak) x_pa / (1 + (r / r_peak)^2)`` (``model.py:171-180``) and Plummer
dispersion profile ``sigma_los = sigma_max / (1 + r^2 / a^2)^(1/4)`` (``model.py:126-128``).
``ModelFitGB`` (``model.py:338-510``) adds the fitted Gaussian background,
``ModelFitConstantBackground`` (``model.py:513-687``) a fixed background likelihood column with the
fitted fraction ``f_back``.
"""
import logging
import os

import numpy as np

from .. import _native
from .. import units as u
from ..parameter import Parameters
from .runner import Runner

logger = logging.getLogger(__name__)
_CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config')


class ModelFit(Runner):
    MODEL_PARAMETERS = ['v_sys', 'v_maxx', 'v_maxy', 'r_peak', 'sigma_max', 'a', 'ra_center', 'dec_center']
    OBSERVABLES = {'v': u.km_s, 'verr': u.km_s, 'ra': u.deg, 'dec': u.deg}

    parameters_file = os.path.join(_CONFIG, 'model.json')

    ROTATION = _native.ROT_RADIAL
    BACKGROUND = _native.BG_NONE

    def __init__(self, data, parameters=None, **kwargs):
        self.ra = None
        self.dec = None
        if parameters is None:
            parameters = Parameters().load(self.parameters_file)
        super(ModelFit, self).__init__(data=data, parameters=parameters, **kwargs)


class ModelFitGB(ModelFit):
    """Radial-profile fit plus a Gaussian background population (``model.py:338-456``)."""

    MODEL_PARAMETERS = ModelFit.MODEL_PARAMETERS + ['v_back', 'sigma_back', 'f_back']
    OBSERVABLES = dict(ModelFit.OBSERVABLES, **{'density': u.dimensionless_unscaled})

    parameters_file = os.path.join(_CONFIG, 'model_with_background.json')

    BACKGROUND = _native.BG_GAUSSIAN

    def __init__(self, data, parameters=None, **kwargs):
        self.density = None
        background = kwargs.pop('background', None)
        if background is not None:
            logger.error('Class ModelFitGB does not support additional background components.')
        if parameters is None:
            parameters = Parameters().load(self.parameters_file)
        super(ModelFitGB, self).__init__(data=data, parameters=parameters, **kwargs)


class ModelFitConstantBackground(ModelFit):
    """Radial-profile fit with a fixed background likelihood and fitted background fraction
    (``model.py:513-623``)."""

    MODEL_PARAMETERS = ModelFit.MODEL_PARAMETERS + ['f_back', ]
    OBSERVABLES = dict(ModelFit.OBSERVABLES, **{'density': u.dimensionless_unscaled})

    parameters_file = os.path.join(_CONFIG, 'model_with_background.json')

    BACKGROUND = _native.BG_FIXED_DENSITY

    def __init__(self, data, background, parameters=None, **kwargs):
        self.density = None
        if parameters is None:
            parameters = Parameters().load(self.parameters_file)
        super(ModelFitConstantBackground, self).__init__(data=data, parameters=parameters, **kwargs)
        self.background = background
        self.lnlike_background = self.background(self.v, self.verr)      # model.py:562-563

    def _background_mode(self):
        return _native.BG_FIXED_DENSITY

    def lnlike(self, values, no_sum=False):
        """``model.py:565-623``; ``no_sum=True`` returns the per-star log-likelihoods of ONE parameter
        vector (``model.py:620-621``)."""
        if not no_sum:
            return super(ModelFitConstantBackground, self).lnlike(values)
        values = np.asarray(values, dtype=np.float64)
        assert values.ndim == 1 and values.size == self.n_fitted_parameters, 'Not all parameters used.'
        return self.pack().lnlike_per_star(values)
