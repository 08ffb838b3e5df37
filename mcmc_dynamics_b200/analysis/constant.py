"""Constant rotation + dispersion fits (``mcmc_dynamics/analysis/constant.py``).

``ConstantFit`` (``constant.py:18-214``): ``v_los = v_sys + v_max sin(theta_i - theta_0)``,
``sigma_los = sigma_max``.  ``ConstantFitGB`` (``constant.py:250-374``) adds a fitted Gaussian
background in velocity and the density prior ``m = density / (density + f_back)``.
The arithmetic lives in ``csrc/mcd_kernels.cu``; these classes only choose the kernel variant.
"""
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
