"""Per-radial-bin fits as ONE batched problem.

The reference's real workflow fits a ``ConstantFit`` to every radial bin in turn
(``bin/run.py:179-190``: 16 walkers x 300 steps per bin; ``bin/run_tests.py:81-97``: 100 walkers x
100 steps per bin), each an independent, latency-bound MCMC over 50-500 stars.  ``RadialBinsFit``
packs all bins into one segmented device handle: one launch evaluates every walker of every bin, and
the device sampler advances all per-bin ensembles together (``csrc/mcd_kernels.cu`` segments,
``csrc/mcd_sampler.cu``).  Bins share the model class, the ``Parameters`` object (bounds, fixed
values, initials) and the optional background object, exactly like the loop bodies of the scripts.
"""
import logging

import numpy as np

from .. import _native
from .. import pack
from .. import sampler as _sampler
from .. import units as u
from ..data_reader import DataReader
from .constant import ConstantFit

logger = logging.getLogger(__name__)


class RadialBinsFit(object):
    """Fit `model_class` independently to every radial bin of `data` (column ``bin``, as written by
    ``DataReader.make_radial_bins``), all bins in one launch.

    Parameters
    ----------
    data : DataReader with a ``bin`` column (integers 0..B-1)
    model_class : a model class without fitted background parameters (default ``ConstantFit``)
    parameters : Parameters shared by all bins, or None for the class default
    background : None, or the background object of ``ConstantFit(data_i, parameters=parameters,
        background=background)`` (``bin/run.py:186``): every bin is then fitted with the fixed-background
        mixture, which needs the ``pmember`` column (``analysis/runner.py:272-286``)
    """

    def __init__(self, data, model_class=ConstantFit, parameters=None, background=None, device=0, math_mode='fast'):
        assert isinstance(data, DataReader), "'data' must be instance of {0}".format(DataReader.__module__)
        if 'bin' not in data.data.columns:
            raise IOError("Input data missing required column <bin>; call make_radial_bins() first.")
        if model_class.BACKGROUND != _native.BG_NONE:
            raise NotImplementedError('RadialBinsFit supports the models without fitted background parameters')
        labels = np.asarray(getattr(data.data['bin'], 'value', data.data['bin'])).astype(np.int64)
        if labels.min() < 0:
            raise ValueError('negative bin labels: every star must belong to a bin')
        self.n_bins = int(labels.max()) + 1
        order = np.argsort(labels, kind='stable')
        counts = np.bincount(labels, minlength=self.n_bins)
        self.segment_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.bin_sizes = counts
        self.order = order
        # a template model on the whole (bin-sorted) catalogue provides validation, units, parameters
        self.background = background
        self.template = model_class(DataReader(data.data[order]), parameters=parameters, background=background,
                                    device=device, math_mode=math_mode)
        self.parameters = self.template.parameters
        self.model_class = model_class
        self.device = device
        self._packed = None
        self._signature = None

    @property
    def fitted_parameters(self):
        return self.template.fitted_parameters

    @property
    def n_fitted_parameters(self):
        return self.template.n_fitted_parameters

    def pack(self):
        t = self.template
        signature = (pack.routing_signature(t.parameters, t.MODEL_PARAMETERS), t.math_mode)
        if self._packed is not None and signature == self._signature:
            return self._packed
        derived = pack.derived_parameters(t.parameters)
        if derived:
            # the batched bins run on the device sampler, which cannot evaluate host-side expressions
            raise pack.PackError(
                "Parameter(s) {0} are constrained by expressions that depend on sampled parameters; per-walker "
                "constraint expressions are not supported by the batched radial-bin fit (fit the bins with "
                "individual models and sampler='host' instead).".format(derived))
        desc, keep = pack.build_descriptor(
            t.parameters, t.MODEL_PARAMETERS, rotation=t.ROTATION, background=t._background_mode(),
            columns=t._star_columns() if self._packed is None else {},
            math_mode=_native.MATH_FAST if t.math_mode == 'fast' else _native.MATH_PLAIN, device=self.device,
            segment_offsets=self.segment_offsets)
        if self._packed is None:
            self._packed = pack.PackedModel(desc, keep)
        else:
            desc.n_stars = self._packed.n_stars
            self._packed.reconfigure(desc)
        self._signature = signature
        return self._packed

    def _theta(self, values):
        theta = np.asarray(values, dtype=np.float64)
        assert theta.ndim == 3 and theta.shape[0] == self.n_bins and theta.shape[2] == self.n_fitted_parameters, \
            'theta must have shape (n_bins, n_walkers, n_fitted_parameters)'
        return theta

    def lnprob(self, values):
        """``[n_bins, n_walkers, n_fitted] -> [n_bins, n_walkers]``: ``Runner.lnprob`` of every bin's model."""
        return self.pack().lnprob(self._theta(values))

    def lnlike(self, values):
        return self.pack().lnlike(self._theta(values))

    def lnprior(self, values):
        theta = self._theta(values)
        return np.stack([self.template._lnprior_batch(theta[b]) for b in range(self.n_bins)])

    def get_initials(self, n_walkers):
        """Independent start positions per bin: ``[n_bins, n_walkers, n_fitted]`` (runner.py:308-330)."""
        return np.stack([self.template.get_initials(n_walkers) for _ in range(self.n_bins)])

    def bin_model(self, i):
        """A stand-alone model object on bin `i` (the reference's loop body), e.g. for cross-checks."""
        lo, hi = self.segment_offsets[i], self.segment_offsets[i + 1]
        sub = DataReader(self.template.data.data[np.arange(lo, hi)])
        return self.model_class(sub, parameters=self.parameters.copy(), background=self.background,
                                device=self.device, math_mode=self.template.math_mode)

    def __call__(self, n_walkers=100, n_steps=100, pos=None, seed=None):
        """Run every bin's ensemble (``cf(n_walkers=100, n_steps=100)`` of ``bin/run_tests.py:97``) on
        the device; returns the sampler, whose ``chain`` is ``[n_bins, n_walkers, n_steps, n_fitted]``."""
        if pos is None:
            pos = self.get_initials(n_walkers)
        pos = np.asarray(pos, dtype=np.float64)
        assert pos.shape == (self.n_bins, n_walkers, self.n_fitted_parameters), \
            'Array with starting values has invalid shape.'
        prior = self.lnprior(pos)
        if not np.all(np.isfinite(prior)):
            b, w = np.argwhere(~np.isfinite(prior))[0]
            raise ValueError("Invalid initial guesses for walker {0} of bin {1}: {2}={3}".format(
                w, b, self.fitted_parameters, pos[b, w]))
        engine = _sampler.DeviceEnsembleSampler(n_walkers, self.n_fitted_parameters, self.pack(), seed=seed)
        engine.run_mcmc(pos, n_steps)
        return engine
