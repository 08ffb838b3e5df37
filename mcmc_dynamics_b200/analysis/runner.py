"""``Runner``: parent of the kinematic model classes, B200 edition.

Same public protocol as ``mcmc_dynamics/analysis/runner.py:23-443`` -- ``lnprior / lnlike /
lnprob(values)``, ``get_initials``, ``fetch_parameter_values``, ``fitted_parameters``, ``units``,
``labels``, ``__call__(n_walkers, n_steps, ...)`` -- but every likelihood evaluation is one CUDA
launch over all walkers handed in.  ``values`` may be the reference's 1-D vector of the free
parameters (a float comes back) or an ``[n_walkers, n_free]`` array, which is what emcee passes with
``vectorize=True`` (an ``[n_walkers]`` float64 array comes back).

There is no CPU implementation of the likelihood in this package.
"""
import logging
import pickle
import warnings

import numpy as np

from .. import _native
from .. import expressions
from .. import pack
from .. import sampler as _sampler
from .. import units as u
from ..background import Gaussian, SingleStars
from ..data_reader import DataReader
from ..parameter import Parameters

logger = logging.getLogger(__name__)


class Runner(object):
    """Parent class of the analysis classes (``analysis/runner.py:23-34``).  Subclasses state their
    ``MODEL_PARAMETERS``, ``OBSERVABLES``, default ``parameters_file`` and which kernel variant
    (``ROTATION``, ``BACKGROUND``) evaluates their ``lnlike``."""

    MODEL_PARAMETERS = []
    OBSERVABLES = {'v': u.km_s, 'verr': u.km_s}
    parameters_file = None

    #: kernel variant (see include/mcd_b200.h)
    ROTATION = _native.ROT_CONSTANT
    BACKGROUND = _native.BG_NONE

    def __init__(self, data, parameters, seed=123, background=None, device=0, math_mode='fast', **kwargs):
        # same checks, in the same order, as analysis/runner.py:55-106
        assert not kwargs, "Unknown keyword arguments provided: {0}".format(kwargs)
        np.random.seed(seed)

        self.v = None
        self.verr = None

        assert isinstance(data, DataReader), "'data' must be instance of {0}".format(DataReader.__module__)
        self.data = data

        if 'ra' in self.OBSERVABLES or 'dec' in self.OBSERVABLES:
            if not data.has_coordinates:
                raise IOError('Missing WCS coordinates of observed data.')

        for required, unit in self.OBSERVABLES.items():
            assert required in data.data.columns, "Input data missing required column <{0}>".format(required)
            quantity = u.as_quantity(data.data[required])
            if quantity.unit.is_unity() and not unit.is_unity():
                quantity = u.Quantity(quantity.value, unit)
                logger.warning('Missing units for <{0}> values. Assuming {1}.'.format(required, unit))
            setattr(self, required, quantity)

        assert isinstance(parameters, Parameters), "'parameters' must be instance of {0}".format(
            Parameters.__module__)
        self.parameters = parameters

        missing = set(self.MODEL_PARAMETERS).difference(self.parameters)
        if missing:
            raise IOError("Missing required parameter(s): '{0}'".format(missing))

        unused = set(self.parameters).difference(self.MODEL_PARAMETERS)
        if unused:
            logger.warning("Superfluous parameter(s) provided: '{0}'".format(unused))

        self.background = background
        if self.background:
            assert isinstance(background, (SingleStars, Gaussian)), \
                "'background' must be an instance of a Background class."
            if 'pmember' not in self.data.data.columns:
                logger.error('Inclusion of background population requires prior probabilities for membership.')
            self.lnlike_background = self.background(self.v, self.verr)
            self.pmember = data.data['pmember']
        else:
            self.lnlike_background = None
            self.pmember = None

        self.device = int(device)
        if math_mode not in ('fast', 'plain'):
            raise ValueError("math_mode must be 'fast' or 'plain'")
        self.math_mode = math_mode
        self._packed = None
        self._packed_signature = None
        self._packed_stamps = None
        self._expression_priors_present = False
        self._derived = []
        self._n_free = None

    # ------------------------------------------------------------------------------------------
    # introspection (analysis/runner.py:108-141, 662-673)
    # ------------------------------------------------------------------------------------------
    @classmethod
    def default_parameters(cls):
        if cls.parameters_file is None:
            raise NotImplementedError
        return Parameters().load(open(cls.parameters_file))

    @property
    def n_data(self):
        return self.data.sample_size

    @property
    def fitted_parameters(self):
        return [p for p in self.parameters if not self.parameters[p].fixed]

    @property
    def n_fitted_parameters(self):
        return len(self.fitted_parameters)

    @property
    def units(self):
        return {p: self.parameters[p].unit for p in self.parameters}

    @property
    def labels(self):
        labels = {}
        for name, parameter in self.parameters.items():
            labels[name] = parameter.label
        return labels

    # ------------------------------------------------------------------------------------------
    # packing
    # ------------------------------------------------------------------------------------------
    def _background_mode(self):
        """Kernel background variant: a plain class switches to the fixed-background mixture when a
        background object was supplied (analysis/runner.py:272-286)."""
        if self.BACKGROUND == _native.BG_NONE and self.background is not None:
            return _native.BG_FIXED_PMEMBER
        return self.BACKGROUND

    def _star_columns(self):
        columns = {
            'ra': u.strip(self.ra, u.deg), 'dec': u.strip(self.dec, u.deg),
            'v': u.strip(self.v, u.km_s), 'verr': u.strip(self.verr, u.km_s),
        }
        mode = self._background_mode()
        if mode == _native.BG_FIXED_PMEMBER:
            columns['pmember'] = u.strip(self.pmember, None)
        if mode in (_native.BG_FIXED_DENSITY, _native.BG_GAUSSIAN):
            columns['density'] = u.strip(getattr(self, 'density'), None)
        if mode in (_native.BG_FIXED_PMEMBER, _native.BG_FIXED_DENSITY):
            columns['lnlike_background'] = u.strip(self.lnlike_background, None)
        return columns

    def _descriptor(self):
        self._derived = pack.derived_parameters(self.parameters)
        return pack.build_descriptor(
            self.parameters, self.MODEL_PARAMETERS, rotation=self.ROTATION, background=self._background_mode(),
            columns=self._star_columns() if self._packed is None else {},
            math_mode=_native.MATH_FAST if self.math_mode == 'fast' else _native.MATH_PLAIN, device=self.device,
            derived=self._derived)

    def pack(self):
        """Upload the star columns (first call) and compile the current parameter routing.  Called
        lazily by every likelihood entry point; cheap when nothing changed."""
        stamps = (pack.edit_stamps(self.parameters), self.math_mode)
        if self._packed is not None and stamps == self._packed_stamps:
            return self._packed                       # nothing was assigned to any parameter since the last call
        self._n_free = self.n_fitted_parameters
        signature = (pack.routing_signature(self.parameters, self.MODEL_PARAMETERS), self.math_mode)
        self._expression_priors_present = any(par.lnprior is not None for par in self.parameters.values())
        self._derived = pack.derived_parameters(self.parameters)
        if self._packed is not None and signature == self._packed_signature:
            self._packed_stamps = stamps
            return self._packed
        desc, keep = self._descriptor()
        if self._packed is None:
            self._packed = pack.PackedModel(desc, keep)
        else:
            desc.n_stars = self._packed.n_stars
            self._packed.reconfigure(desc)
        self._packed_signature = signature
        self._packed_stamps = stamps
        return self._packed

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_packed'] = None              # device handles do not pickle; re-packed on first use
        state['_packed_signature'] = None
        state['_packed_stamps'] = None
        return state

    # ------------------------------------------------------------------------------------------
    # parameters and priors
    # ------------------------------------------------------------------------------------------
    def fetch_parameter_values(self, values):
        """``analysis/runner.py:143-180``: dictionary name -> quantity for one parameter vector, fixed
        parameters merged in.  Unlike the reference it does not write the values back into
        ``self.parameters`` (a side effect no caller relies on, and a hazard for batched calls)."""
        current_parameters = {}
        i = 0
        for name, parameter in self.parameters.items():
            if parameter.fixed:
                v = u.Quantity(parameter.value, parameter.unit)
            else:
                v = u.Quantity(values[i], parameter.unit)
                i += 1
            current_parameters[name] = v
        assert i == len(values), 'Not all parameters used.'
        return current_parameters

    def _as_batch(self, values, packed=False):
        """One vector or a batch -> ``([n, n_free] float64, was_one_vector)``.  ``packed=True``: ``pack()`` has
        just run, so its cached count of sampled parameters is current (no second look at the edit stamps)."""
        values = np.asarray(values, dtype=np.float64)
        scalar = values.ndim == 1
        theta = values[None, :] if scalar else values
        n_free = self._n_free if packed else self.n_fitted_parameters
        assert theta.ndim == 2 and theta.shape[1] == n_free, 'Not all parameters used.'
        return theta, scalar

    def _derived_columns(self, theta):
        """Per-walker values of the ``expr``-constrained parameters that depend on sampled ones
        (``analysis/runner.py:163-176`` evaluates them through the shared asteval table on every call,
        ``parameter.py:865-874``), shape ``[n_walkers, n_derived]`` in ``self._derived`` order.  Every
        sampled parameter carries the walker's value when an expression is evaluated; the reference's
        iteration-order staleness (a parameter later in the table still holds the previous call's value
        during ``lnprior``) is not reproduced.  Nothing is written back into ``self.parameters``."""
        names = self._derived
        n = theta.shape[0]
        out = np.empty((n, len(names)), dtype=np.float64)
        if not names or n == 0:
            return out
        self.parameters._sync_symbols()
        base = dict(self.parameters.symtable)
        free = self.fitted_parameters
        trees = [self.parameters[name]._expr_ast for name in names]

        def one_row(w):
            sym = dict(base)
            for j, name in enumerate(free):
                sym[name] = float(theta[w, j])
            row = np.empty(len(names))
            for k, name in enumerate(names):
                row[k] = float(expressions.evaluate(trees[k], sym))
                sym[name] = row[k]
            return row

        first = one_row(0)
        try:                                   # all rows at once where the expressions are plain arithmetic
            sym = dict(base)
            for j, name in enumerate(free):
                sym[name] = theta[:, j]
            for k, name in enumerate(names):
                column = np.asarray(expressions.evaluate(trees[k], sym), dtype=np.float64)
                out[:, k] = np.broadcast_to(column, (n,))
                sym[name] = out[:, k]
            if np.array_equal(out[0], first, equal_nan=True):
                return out
        except Exception:                      # noqa: BLE001 -- e.g. builtin min/max or a conditional on arrays
            pass
        out[0] = first
        for w in range(1, n):
            out[w] = one_row(w)
        return out

    def _device_theta(self, theta):
        """theta as the kernel wants it: the free parameters followed by the per-walker constrained ones."""
        if not self._derived:
            return theta
        return np.ascontiguousarray(np.hstack([theta, self._derived_columns(theta)]))

    def _lnprior_batch(self, theta, parameters_to_ignore=None):
        """Box prior over every parameter plus optional expression priors (runner.py:206-217,
        parameter.py:684-705), for all rows of theta at once."""
        lnp = np.zeros(theta.shape[0], dtype=np.float64)
        derived = pack.derived_parameters(self.parameters)
        derived_values = None
        if derived:
            self._derived = derived
            derived_values = self._derived_columns(theta)
        j = 0
        for name, par in self.parameters.items():
            if par.fixed and name not in derived:
                value = par.value
                if value < par.min or value > par.max:
                    lnp[:] = -np.inf
                elif par.lnprior is not None:
                    lnp += par.evaluate_lnprior(value)
            else:
                if par.fixed:
                    column = derived_values[:, derived.index(name)]
                else:
                    column = theta[:, j]
                    j += 1
                outside = (column < par.min) | (column > par.max)
                lnp[outside] = -np.inf
                if par.lnprior is not None:
                    for w in np.flatnonzero(~outside & np.isfinite(lnp)):
                        lnp[w] += par.evaluate_lnprior(column[w])
        lnp[~np.isfinite(lnp)] = -np.inf
        return lnp

    def lnprior(self, values, parameters_to_ignore=None):
        """``analysis/runner.py:182-217``; accepts one vector or a batch."""
        theta, scalar = self._as_batch(values)
        lnp = self._lnprior_batch(theta, parameters_to_ignore)
        if scalar:
            return -np.inf if not np.isfinite(lnp[0]) else (0 if lnp[0] == 0 else lnp[0])
        return lnp

    # ------------------------------------------------------------------------------------------
    # likelihood: one CUDA launch per call
    # ------------------------------------------------------------------------------------------
    def lnlike(self, values):
        """Log-likelihood without priors (``constant.py:113-154``, ``model.py:182-223`` and the
        background variants), for one parameter vector or a batch."""
        packed = self.pack()
        theta, scalar = self._as_batch(values, packed=True)
        out = packed.lnlike(self._device_theta(theta))
        return float(out[0]) if scalar else out

    def lnprob(self, values):
        """``analysis/runner.py:288-306``: box prior fused into the kernel; walkers outside the prior
        come back as exactly ``-inf`` without being evaluated."""
        packed = self.pack()
        theta, scalar = self._as_batch(values, packed=True)
        out = packed.lnprob(self._device_theta(theta))
        if self._expression_priors_present:           # refreshed by pack() whenever a parameter was edited
            extra = self._lnprior_batch(theta)
            with np.errstate(invalid='ignore'):
                out = np.where(np.isfinite(extra), out + extra, -np.inf)
        return float(out[0]) if scalar else out

    def _model_curves(self, who, given):
        """``(v_los, sigma_los)`` of every star in km/s for the parameter values handed to ``rotation_model`` /
        ``dispersion_model``: one launch of the per-star kernel (``mcd_model_per_star``).  `given` maps
        parameter names to quantities or plain numbers (taken to be in the parameter's own unit, as the
        reference takes them); sampled parameters that the curve does not depend on keep their current value.
        A parameter that is FIXED in this model lives in the packed columns (a fixed centre is folded into the
        stored tangent-plane coordinates), so a value that differs from it cannot be honoured."""
        packed = self.pack()
        theta = np.empty(self._n_free, dtype=np.float64)
        j = 0
        for name, par in self.parameters.items():
            value = given.get(name)
            if value is not None:
                value = float(u.strip(value, par.unit))
            if par.fixed:
                if name in self._derived or value is None or name not in self.MODEL_PARAMETERS:
                    continue
                if abs(value - par.value) > 1e-12 * max(1.0, abs(par.value)):
                    raise ValueError(
                        "{0}.{1}: '{2}' is fixed at {3} in this model and part of the packed star columns; free the "
                        "parameter (or change its value in .parameters) to evaluate the model at {4}".format(
                            self.__class__.__name__, who, name, par.value, value))
            else:
                if value is None:           # the requested curve does not depend on it: any sane number
                    value = float(par.value) if par.value not in (None, 0) else 1.0
                theta[j] = value
                j += 1
        v_los, sigma_los = packed.model_per_star(self._device_theta(theta[None, :])[0])
        return u.Quantity(v_los, u.km_s), u.Quantity(sigma_los, u.km_s)

    @staticmethod
    def _no_kwargs(cls_name, who, kwargs):
        if kwargs:              # constant.py:70-72,102-104; model.py:121-123,164-166
            raise IOError('Unknown keyword argument(s) "{0}" for method {1}.{2}.'.format(
                ', '.join(kwargs.keys()), cls_name, who))

    def _calculate_lnlike(self, v_los, sigma_los):
        """``analysis/runner.py:240-286``: the log-likelihood of the data for model curves computed by the caller
        (the hook a user-defined model class builds its ``lnlike`` on): Gaussian sum, or the mixture with the
        fixed background column when a ``background=`` object was given.  `v_los`, `sigma_los`: one value per
        star, quantities or plain numbers in km/s.  One reduction launch over the resident ``v`` / ``verr``
        (/ ``pmember`` / ``lnlike_background``) columns; the model classes of this package never call it --
        their curves are evaluated inside the fused kernel."""
        packed = self.pack()
        return packed.calculate_lnlike(u.strip(v_los, u.km_s), u.strip(sigma_los, u.km_s))

    def _has_expression_priors(self):
        return any(par.lnprior is not None for par in self.parameters.values())

    def lnprob_tensor(self, theta):
        """``lnprob`` for an ``[n_walkers, n_free]`` float64 CUDA tensor, asynchronous on the current
        stream, result stays on the device.  Box priors only: expression priors (``lnprior`` strings) and
        per-walker ``expr`` constraints are evaluated on the host, so a model that has either raises
        here instead of silently returning a different posterior -- use :meth:`lnprob`."""
        packed = self.pack()
        self._require_box_priors('lnprob_tensor')
        return packed.lnprob_tensor(theta)

    def _require_box_priors(self, what):
        if self._expression_priors_present or self._derived:
            raise ValueError(
                "{0} supports box priors only: this model has {1}, which are evaluated on the host "
                "(use lnprob / sampler='host')".format(
                    what, 'expression priors' if self._expression_priors_present
                    else "per-walker 'expr' constraints ({0})".format(', '.join(self._derived))))

    # ------------------------------------------------------------------------------------------
    # sampling
    # ------------------------------------------------------------------------------------------
    def get_initials(self, n_walkers):
        """``analysis/runner.py:308-330``."""
        initials = np.zeros((n_walkers, self.n_fitted_parameters))
        i = 0
        for name, parameter in self.parameters.items():
            if parameter.fixed:
                continue
            else:
                initials[:, i] = parameter.evaluate_initials(n_walkers)
            i += 1
        return initials

    def __call__(self, n_walkers=100, n_steps=500, n_burn=100, n_threads=1, n_out=None, pos=None, lnprob0=None,
                 plot=False, prefix='sampler', true_values=None, sampler='auto', seed=None, **kwargs):
        """Run the ensemble sampler (``analysis/runner.py:332-443``).

        ``n_threads`` is accepted and ignored: the walker-parallel process pool of the reference is
        replaced by the batched launch.  ``sampler='host'`` runs an emcee-compatible stretch-move
        loop on the host that calls ``lnprob`` in vectorised mode (emcee itself is used when it is
        importable); ``sampler='device'`` keeps the whole chain on the GPU (same move, counter-based
        random numbers).  ``sampler='auto'`` (default) takes the device engine whenever it can express
        the model -- box priors only, no per-walker ``expr`` constraints, no ``lnprob0`` handed in -- and
        the host loop otherwise: on the small configurations a host round trip per half-step costs
        10-20x the likelihood itself (profiles/r02_configs.md), so the drop-in default should not pay it.
        """
        if kwargs:
            if "filename" in kwargs or "plotfilename" in kwargs:
                logger.warning('Parameters <filename> and <plotfilename> not used anymore. Use <prefix> instead.')
        if plot:
            logger.warning('Plotting is outside the scope of this package; plot=True is ignored.')

        if pos is not None:
            pos = np.asarray(pos, dtype=np.float64)
            assert pos.shape == (n_walkers, self.n_fitted_parameters), 'Array with starting values has invalid shape.'
        else:
            pos = self.get_initials(n_walkers=n_walkers)

        prior = self._lnprior_batch(pos)
        for i in range(n_walkers):
            if not np.isfinite(prior[i]):
                raise ValueError(
                    "Invalid initial guesses for walker {0}: {1}={2}".format(i, self.fitted_parameters, pos[i]))

        if sampler not in ('auto', 'device', 'host'):
            raise ValueError("sampler must be 'auto', 'device' or 'host'")
        if sampler == 'auto':
            self.pack()
            box_only = not (self._expression_priors_present or self._derived)
            sampler = 'device' if (box_only and lnprob0 is None and n_walkers <= 4096) else 'host'
        if sampler == 'device':
            packed = self.pack()
            self._require_box_priors("sampler='device'")
            engine = _sampler.DeviceEnsembleSampler(n_walkers, self.n_fitted_parameters, packed, seed=seed)
        else:
            engine = _sampler.make_host_sampler(n_walkers, self.n_fitted_parameters, self.lnprob, seed=seed)
        logger.info("Running MCMC chain ...")

        if n_out is not None:
            msg = "Iter. <log like>   "
            for name, parameter in self.parameters.items():
                if not parameter.fixed:
                    msg += " {0:12s}".format('<' + name + '>')
            logger.info(msg)

        state = None
        while engine.iteration < n_steps:
            todo = n_out if n_out is not None else n_steps
            todo = min(todo, n_steps - engine.iteration)
            pos, lnp, state = engine.run_mcmc(pos, todo, log_prob0=lnprob0, rstate0=state, progress=False)
            lnprob0 = None
            if n_out is not None:
                output = " {0:4d} {1:12.5e}".format(engine.iteration, np.mean(lnp[:]))
                for i in range(self.n_fitted_parameters):
                    output += " {0:12.5e}".format(np.mean(pos[:, i]))
                if engine.iteration % n_out == 0 and prefix is not None:
                    self.save_current_status(engine, prefix=prefix)
                logger.info(output)
        return engine

    @staticmethod
    def save_current_status(sampler, prefix="sampler"):
        """``analysis/runner.py:458-477``: chain and log-probabilities pickled to two files."""
        with open("{0}_chain.pkl".format(prefix), "wb") as f:
            pickle.dump(sampler.chain, f)
        with open("{0}_lnprob.pkl".format(prefix), "wb") as f:
            pickle.dump(sampler.lnprobability, f)

    @staticmethod
    def save_chain(sampler, filename="samplerchain.pkl"):
        """Deprecated in the reference too (``analysis/runner.py:445-455``): forwards to
        :meth:`save_current_status` with the prefix derived from `filename`."""
        warnings.warn('Method Runner.save_chain() is deprecated. Use Runner.save_current_status() instead.',
                      DeprecationWarning)
        prefix = filename.split('.')[0]
        if len(prefix) > 5 and prefix[-5:] == 'chain':
            prefix = prefix[:-5]
        Runner.save_current_status(sampler, prefix=prefix)

    def sample_chain(self, chain, n_burn, n_samples=1):
        """``analysis/runner.py:820-850``: `n_samples` parameter sets drawn at random (NumPy's global generator,
        seeded by the constructor's `seed`) from the post-burn-in part of `chain`, each as the dictionary
        :meth:`fetch_parameter_values` returns."""
        chain = np.asarray(chain)
        _parameters = np.reshape(chain[:, n_burn:], (-1, chain.shape[-1]))
        indices = np.random.randint(0, _parameters.shape[0], (n_samples,))
        return [self.fetch_parameter_values(row) for row in _parameters[indices]]

    @staticmethod
    def read_chain(filename):
        """``analysis/runner.py:480-496``."""
        with open(filename, 'rb') as f:
            return pickle.load(f)

    @staticmethod
    def read_final_chain(filename):
        """Last positions of every walker, for resuming a run (``analysis/runner.py:499-519``)."""
        chain = Runner.read_chain(filename)
        return chain[:, -1, :]

    def convert_to_parameters(self, chain, n_burn):
        """``analysis/runner.py:521-564``: one array of post-burn-in samples per parameter -- sampled
        ones from the chain, fixed ones repeated, ``expr``-constrained ones evaluated per sample."""
        chain = np.asarray(chain)
        pars = {}
        n_samples = chain.shape[0] * (chain.shape[1] - n_burn)
        for par in self.parameters:
            if par in self.fitted_parameters:
                i = self.fitted_parameters.index(par)
                pars[par] = chain[:, n_burn:, i].flatten()
        for fix_par in [p for p in self.parameters if p not in pars]:
            if self.parameters[fix_par].expr is None:
                pars[fix_par] = np.full(n_samples, self.parameters[fix_par].value)
        for dep_par in [p for p in self.parameters if p not in pars]:
            if self.parameters[dep_par].expr is not None:
                values = np.zeros(n_samples, dtype=np.float64)
                deps = [p for p in pars.keys() if p in self.parameters[dep_par]._expr_deps]
                for n in range(n_samples):
                    for par in deps:
                        self.parameters[par].value = pars[par][n]
                    values[n] = self.parameters[dep_par].value
                pars[dep_par] = values
        return pars

    def compute_theta_vmax(self, chain, n_burn, return_samples=False):
        """Position angle ``theta_0`` and amplitude ``v_max`` of the rotation field from the
        ``(v_maxx, v_maxy)`` samples (``constant.py:156-214``, ``model.py:319-335``,
        ``utils/coordinates/get_amplitude_and_angle.py:10-51``): median and 16/84 percentiles."""
        pars = self.convert_to_parameters(chain=chain, n_burn=n_burn)
        results, v_max, _theta = get_amplitude_and_angle(pars, return_samples=return_samples)
        if results is None:
            logger.error('Could not recover paramaters of rotation field in {}.compute_theta_vmax().'.format(
                self.__class__.__name__))
            return None
        results.units['v_max'] = self.units['v_maxx']
        if return_samples:
            return results, v_max, _theta, pars['sigma_max']
        return results

    def compute_percentiles(self, chain, n_burn, pct=None):
        """``analysis/runner.py:566-613``: the requested percentiles (default 16, 50, 84) of the
        post-burn-in samples of every fitted parameter, shape ``[len(pct), n_fitted]``."""
        if pct is None:
            pct = [16, 50, 84]
        _samples = np.asarray(chain)[:, n_burn:, :].reshape((-1, self.n_fitted_parameters))
        return np.percentile(_samples, pct, axis=0)

    def compute_bestfit_values(self, chain, n_burn):
        """``analysis/runner.py:615-660``: median and upper / lower uncertainty of every fitted
        parameter.  The reference returns an indexed astropy table; this returns a
        :class:`BestFit` with the same access pattern (``bestfit.columns``,
        ``bestfit.loc['median'][name]``) and, like the reference, stores the medians as the
        parameters' current values."""
        percentiles = self.compute_percentiles(chain, n_burn=n_burn, pct=[16, 50, 84])
        columns = {}
        i = 0
        for name, parameter in self.parameters.items():
            if parameter.fixed:
                continue
            parameter.value = percentiles[1, i]
            columns[name] = (percentiles[1, i], percentiles[2, i] - percentiles[1, i],
                             percentiles[1, i] - percentiles[0, i])
            i += 1
        return BestFit(columns, {name: self.parameters[name].unit for name in columns})

    def calculate_membership_probabilities(self, chain, n_burn):
        """A-posteriori cluster membership probability of every star at the posterior median
        (``constant.py:366-374``, ``model.py:458-510,625-687``): one per-star kernel launch.
        Available for every model with a background component."""
        if self._background_mode() == _native.BG_NONE:
            raise NotImplementedError('membership probabilities need a background component')
        median = self.compute_percentiles(chain, n_burn=n_burn, pct=[50])[0]
        self.compute_bestfit_values(chain, n_burn)          # the reference updates parameter values here
        packed = self.pack()
        return packed.membership_per_star(self._device_theta(np.atleast_2d(median))[0])


def get_amplitude_and_angle(pars, return_samples=False):
    """``utils/coordinates/get_amplitude_and_angle.py:10-51``: complete the triple (theta_0, v_maxx,
    v_maxy) from whichever two are given, measure angles relative to the direction of the median
    velocity vector (so that the median sits in the middle of (-pi, pi]), and define ``v_max`` as the
    component of (v_maxx, v_maxy) along that direction.  Returns ``(BestFit, v_max, theta)``."""
    pars = dict(pars)
    if 'theta_0' not in pars and 'v_maxx' in pars and 'v_maxy' in pars:
        pars['theta_0'] = np.arctan2(pars['v_maxy'], pars['v_maxx'])
    elif 'v_maxx' not in pars and 'theta_0' in pars and 'v_maxy' in pars:
        pars['v_maxx'] = pars['v_maxy'] * np.tan(pars['theta_0'])
    elif 'v_maxy' not in pars and 'theta_0' in pars and 'v_maxx' in pars:
        pars['v_maxy'] = pars['v_maxx'] / np.tan(pars['theta_0'])
    for par in ['theta_0', 'v_maxx', 'v_maxy']:
        if par not in pars:
            logger.error('Failed to recover parameter {}.'.format(par))
            return None, None, None

    median_theta = np.arctan2(np.median(pars['v_maxy']), np.median(pars['v_maxx']))
    _theta = pars['theta_0'] - median_theta
    _theta = np.where(_theta < -np.pi, _theta + 2 * np.pi, _theta)
    _theta = np.where(_theta > np.pi, _theta - 2 * np.pi, _theta)
    v_max = pars['v_maxx'] * np.cos(-median_theta) - pars['v_maxy'] * np.sin(-median_theta)

    columns = {}
    for name, values in (('v_max', v_max), ('theta_0', _theta)):
        p = np.percentile(values, [16, 50, 84])
        columns[name] = [p[1], p[2] - p[1], p[1] - p[0]]
    columns['theta_0'][0] += median_theta
    results = BestFit({k: tuple(v) for k, v in columns.items()}, {'v_max': None, 'theta_0': u.rad})
    if return_samples:
        return results, v_max, _theta
    return results, None, None


class BestFit(object):
    """Rows ``median / uperr / loerr`` by parameter (stand-in for the indexed table of
    ``analysis/runner.py:640-660``)."""

    ROWS = ('median', 'uperr', 'loerr')

    def __init__(self, columns, units):
        self._columns = columns
        self.units = units
        self.columns = ['value'] + list(columns)

    @property
    def loc(self):
        return {row: dict([('value', row)] + [(name, u.Quantity(values[k], self.units[name]) if self.units[name]
                                               is not None else values[k])
                                              for name, values in self._columns.items()])
                for k, row in enumerate(self.ROWS)}

    def __getitem__(self, name):
        if name == 'value':
            return list(self.ROWS)
        return np.asarray(self._columns[name])
