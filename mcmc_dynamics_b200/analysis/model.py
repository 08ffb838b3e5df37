"""Radial-profile fits (``mcmc_dynamics/analysis/model.py``).

``ModelFit`` (``model.py:20-335``): Lynden-Bell rotation curve
``v_los = v_sys + 2 (v_max / r_peak) x_pa / (1 + (r / r_peak)^2)`` (``model.py:171-180``) and Plummer
dispersion profile ``sigma_los = sigma_max / (1 + r^2 / a^2)^(1/4)`` (``model.py:126-128``).
``ModelFitGB`` (``model.py:338-510``) adds the fitted Gaussian background,
``ModelFitConstantBackground`` (``model.py:513-687``) a fixed background likelihood column with the
fitted fraction ``f_back``.
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


class ModelFit(Runner):
    MODEL_PARAMETERS = ['v_sys', 'v_maxx', 'v_maxy', 'r_peak', 'sigma_max', 'a', 'ra_center', 'dec_center']
    OBSERVABLES = {'v': u.km_s, 'verr': u.km_s, 'ra': u.deg, 'dec': u.deg}

    parameters_file = config.default_file('model')

    ROTATION = _native.ROT_RADIAL
    BACKGROUND = _native.BG_NONE

    def __init__(self, data, parameters=None, **kwargs):
        self.ra = None
        self.dec = None
        if parameters is None:
            parameters = Parameters().load(self.parameters_file)
        super(ModelFit, self).__init__(data=data, parameters=parameters, **kwargs)
        # which parameters each curve takes (constant.py:49-50, model.py:90-91)
        self.rotation_parameters = inspect.signature(self.rotation_model).parameters
        self.dispersion_parameters = inspect.signature(self.dispersion_model).parameters

    def dispersion_model(self, sigma_max, ra_center, dec_center, a=1, **kwargs):
        """``model.py:93-128``: ``sigma_max / (1 + r^2 / a^2)^(1/4)`` at every star, km/s (GPU, per-star kernel).
        `a` without a unit is taken to be in the unit of the ``a`` parameter."""
        self._no_kwargs(self.__class__.__name__, 'dispersion_model', kwargs)
        return self._model_curves('dispersion_model', {'sigma_max': sigma_max, 'ra_center': ra_center,
                                                       'dec_center': dec_center, 'a': a})[1]

    def rotation_model(self, v_sys, v_maxx, v_maxy, ra_center, dec_center, r_peak=None, **kwargs):
        """``model.py:130-180``: ``v_sys + 2 (v_max / r_peak) x_pa / (1 + (r / r_peak)^2)`` at every star, km/s.
        ``r_peak=None`` takes the median distance of the stars from the centre, as the reference does
        (``model.py:172-173``)."""
        self._no_kwargs(self.__class__.__name__, 'rotation_model', kwargs)
        if r_peak is None:
            r_peak = self.data.compute_distances(ra_center, dec_center)
            r_peak = u.Quantity(np.median(r_peak.value), r_peak.unit)
        return self._model_curves('rotation_model', {'v_sys': v_sys, 'v_maxx': v_maxx, 'v_maxy': v_maxy,
                                                     'ra_center': ra_center, 'dec_center': dec_center,
                                                     'r_peak': r_peak})[0]

    def create_profiles(self, chains, n_burn, radii=None, filename=None):
        """Radial profiles of the rotation amplitude and the velocity dispersion implied by the
        post-burn-in samples (``model.py:225-317``): median, 1-sigma (16/84 %) and 3-sigma
        (0.15/99.85 %) limits at every radius.  Host-side post-processing, O(radii x samples).

        `radii` without unit are taken to be in the unit of ``r_peak``; default
        ``logspace(-1, 2.5, 50)`` arcsec.  Returns a :class:`~mcmc_dynamics_b200.data_reader.Table`.
        """
        from ..data_reader import Table
        chains = np.asarray(chains)
        samples = {}
        i = 0
        for name, parameter in self.parameters.items():
            if parameter.fixed:
                samples[name] = np.asarray(parameter.value, dtype=np.float64)
            else:
                samples[name] = chains[:, n_burn:, i].flatten()
                i += 1
        unit_rp = self.parameters['r_peak'].unit or u.arcsec
        unit_a = self.parameters['a'].unit or u.arcsec
        if radii is None:
            radii = u.Quantity(np.logspace(-1, 2.5, 50), u.arcsec)
        elif not u.is_quantity(radii) or u.as_quantity(radii).unit.is_unity():
            radii = u.Quantity(np.asarray(getattr(radii, 'value', radii), dtype=np.float64), unit_rp)
        radii = u.as_quantity(radii)
        r_in_rp = radii.to(unit_rp).value[:, np.newaxis]
        r_in_a = radii.to(unit_a).value[:, np.newaxis]
        to_kms = (self.parameters['v_maxx'].unit or u.km_s).to(u.km_s)
        v_max = np.sqrt(samples['v_maxx'] ** 2 + samples['v_maxy'] ** 2) * to_kms
        v_rot = 2. * (v_max / samples['r_peak']) * r_in_rp / (1. + (r_in_rp / samples['r_peak']) ** 2)
        sigma_max = samples['sigma_max'] * (self.parameters['sigma_max'].unit or u.km_s).to(u.km_s)
        sigma = sigma_max / (1. + r_in_a ** 2 / samples['a'] ** 2) ** 0.25
        if v_rot.ndim == 1:                     # every parameter fixed: one "sample"
            v_rot, sigma = v_rot[:, np.newaxis], sigma[:, np.newaxis]
        pct = [50, 16, 84, 0.15, 99.85]
        pv_rot = np.percentile(v_rot, pct, axis=-1)
        psigma = np.percentile(sigma, pct, axis=-1)
        profile = Table()
        profile['r'] = radii
        for stem, values in (('v_rot', pv_rot), ('sigma', psigma)):
            for suffix, row in zip(('', '_lower_1s', '_upper_1s', '_lower_3s', '_upper_3s'), values):
                profile[stem + suffix] = u.Quantity(row, u.km_s)
        if filename is not None:
            names = profile.colnames
            np.savetxt(filename, np.column_stack([np.asarray(profile[n].value) for n in names]), delimiter=',',
                       header=','.join(names))
        return profile


class ModelFitGB(ModelFit):
    """Radial-profile fit plus a Gaussian background population (``model.py:338-456``)."""

    MODEL_PARAMETERS = ModelFit.MODEL_PARAMETERS + ['v_back', 'sigma_back', 'f_back']
    OBSERVABLES = dict(ModelFit.OBSERVABLES, **{'density': u.dimensionless_unscaled})

    parameters_file = config.default_file('model_with_background')

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

    parameters_file = config.default_file('model_with_background')

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
        packed = self.pack()
        return packed.lnlike_per_star(self._device_theta(values[None, :])[0])
