"""Constant rotation + dispersion fits (``mcmc_dynamics/analysis/constant.py``).

``ConstantFit`` (``constant.py:18-214``): ``v_los = v_sys + v_max sin(theta_i - theta_0)``,
``sigma_los = sigma_max``.  ``ConstantFitGB`` (``constant.py:250-374``) adds a fitted Gaussian
background in velocity and the density prior ``m = density / (density + f_back)``.
The arithmetic lives in ``csrc/mcd_kernels.cu``; these classes only choose the kernel variant.
"""
import inspect
import logging

import numpy as np

from .. import _native
from .. import config
from .. import units as u
from ..parameter import Parameters
from .runner import Runner

logger = logging.getLogger(__name__)


class ConstantFit(Runner):
    MODEL_PARAMETERS = ['v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center']
    OBSERVABLES = {'v': u.km_s, 'verr': u.km_s, 'ra': u.deg, 'dec': u.deg}

    parameters_file = config.default_file('constant')

    ROTATION = _native.ROT_CONSTANT
    BACKGROUND = _native.BG_NONE

    def __init__(self, data, parameters=None, **kwargs):
        self.ra = None
        self.dec = None
        if parameters is None:
            parameters = Parameters().load(self.parameters_file)
        super(ConstantFit, self).__init__(data=data, parameters=parameters, **kwargs)
        # which parameters each curve takes (constant.py:49-50, model.py:90-91)
        self.rotation_parameters = inspect.signature(self.rotation_model).parameters
        self.dispersion_parameters = inspect.signature(self.dispersion_model).parameters

    def dispersion_model(self, sigma_max, **kwargs):
        """``constant.py:52-74``: the (constant) model dispersion at every star, km/s.  Evaluated on the GPU by
        the per-star kernel, like everything else that touches the star columns."""
        self._no_kwargs(self.__class__.__name__, 'dispersion_model', kwargs)
        return self._model_curves('dispersion_model', {'sigma_max': sigma_max})[1]

    def rotation_model(self, v_sys, v_maxx, v_maxy, ra_center, dec_center, **kwargs):
        """``constant.py:76-111``: ``v_sys + v_max sin(theta_i - theta_0)`` at every star, km/s."""
        self._no_kwargs(self.__class__.__name__, 'rotation_model', kwargs)
        return self._model_curves('rotation_model', {'v_sys': v_sys, 'v_maxx': v_maxx, 'v_maxy': v_maxy,
                                                     'ra_center': ra_center, 'dec_center': dec_center})[0]


class ConstantFitGB(ConstantFit):
    """Constant fit plus a Gaussian background population (``constant.py:250-374``)."""

    MODEL_PARAMETERS = ConstantFit.MODEL_PARAMETERS + ['v_back', 'sigma_back', 'f_back']
    OBSERVABLES = dict(ConstantFit.OBSERVABLES, **{'density': u.dimensionless_unscaled})

    parameters_file = config.default_file('constant_with_background')

    BACKGROUND = _native.BG_GAUSSIAN

    def __init__(self, data, parameters=None, **kwargs):
        self.density = None
        if parameters is None:
            parameters = Parameters().load(self.parameters_file)
        background = kwargs.pop('background', None)
        if background is not None:
            logger.error('Class ConstantFitGB does not support additional background components.')
        super(ConstantFitGB, self).__init__(data=data, parameters=parameters, **kwargs)
