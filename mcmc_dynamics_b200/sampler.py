"""Ensemble samplers behind ``Runner.__call__``.

The reference hands ``self.lnprob`` to ``emcee.EnsembleSampler`` (``analysis/runner.py:403``) and
drives it with ``run_mcmc`` (``runner.py:416-419``).  emcee is an unpinned third-party dependency
that is not vendored by the reference; two stand-ins with the part of its interface the reference
uses (``run_mcmc``, ``chain``, ``lnprobability``, ``iteration``, ``acceptance_fraction``):

* :class:`HostEnsembleSampler` -- the red/blue stretch move (Goodman & Weare 2010; emcee 3
  ``RedBlueMove`` / ``StretchMove(a=2)``) on the host, calling the model's ``lnprob`` once per
  half-ensemble with an ``[n, ndim]`` array (``vectorize=True`` semantics).  If emcee is importable,
  :func:`make_host_sampler` returns a real ``emcee.EnsembleSampler(..., vectorize=True)`` instead.
* :class:`DeviceEnsembleSampler` -- the same move with state, random numbers, proposals and
  accept/reject on the GPU (``csrc/mcd_sampler.cu``), one CUDA graph launch per step.
"""
import ctypes
import math

import numpy as np

from . import _native


class HostEnsembleSampler(object):
    """Minimal emcee-compatible ensemble sampler (vectorised log-probability calls)."""

    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, seed=None, vectorize=True):
        if nwalkers < 2 * ndim:
            raise RuntimeError("It is unadvisable to use a red-blue move with fewer walkers than twice the number "
                               "of dimensions.")
        self.nwalkers = int(nwalkers)
        self.ndim = int(ndim)
        self.log_prob_fn = log_prob_fn
        self.a = float(a)
        self.vectorize = vectorize
        self._random = np.random.RandomState(seed)
        self.iteration = 0
        self._chain = []
        self._lnprob = []
        self.naccepted = np.zeros(self.nwalkers, dtype=np.int64)
        self.n_log_prob_calls = 0

    # -- emcee accessors used by the reference (runner.py:422,429,471-472) -----------------------
    @property
    def chain(self):
        """[nwalkers, nsteps, ndim]"""
        if not self._chain:
            return np.empty((self.nwalkers, 0, self.ndim))
        return np.swapaxes(np.asarray(self._chain), 0, 1)

    @property
    def lnprobability(self):
        """[nwalkers, nsteps]"""
        if not self._lnprob:
            return np.empty((self.nwalkers, 0))
        return np.asarray(self._lnprob).T

    def get_chain(self, discard=0, flat=False):
        chain = np.asarray(self._chain)[discard:]
        return chain.reshape((-1, self.ndim)) if flat else chain

    def get_log_prob(self, discard=0, flat=False):
        lnp = np.asarray(self._lnprob)[discard:]
        return lnp.reshape(-1) if flat else lnp

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(max(1, self.iteration))

    @property
    def random_state(self):
        return self._random.get_state()

    # -- sampling -----------------------------------------------------------------------------
    def compute_log_prob(self, coords):
        coords = np.asarray(coords, dtype=np.float64)
        # one dot product in the common case (a finite sum of squares proves every entry finite), emcee's
        # checks and messages otherwise
        flat = coords.reshape(-1)
        if not math.isfinite(np.dot(flat, flat)) and not np.isfinite(coords).all():
            if np.any(np.isinf(coords)):
                raise ValueError("At least one parameter value was infinite")
            raise ValueError("At least one parameter value was NaN")
        self.n_log_prob_calls += 1
        if self.vectorize:
            log_prob = np.asarray(self.log_prob_fn(coords), dtype=np.float64)
        else:
            log_prob = np.array([float(self.log_prob_fn(row)) for row in coords], dtype=np.float64)
        norm = np.dot(log_prob, log_prob)                 # -inf entries are legal and give +inf; only NaN gives NaN
        if norm != norm and np.isnan(log_prob).any():
            raise ValueError("Probability function returned NaN")
        return log_prob

    def run_mcmc(self, initial_state, nsteps, log_prob0=None, rstate0=None, progress=False, store=True, **kwargs):
        coords = np.array(initial_state, dtype=np.float64)
        if coords.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions {0}".format(coords.shape))
        if rstate0 is not None:
            self._random.set_state(rstate0)
        if self.nwalkers > 1 and np.linalg.cond(np.atleast_2d(np.cov(coords, rowvar=False))) > 1e8:
            raise ValueError("Initial state has a large condition number. Make sure that your walkers are "
                             "linearly independent for the best performance")
        log_prob = self.compute_log_prob(coords) if log_prob0 is None else np.array(log_prob0, dtype=np.float64)
        if np.any(np.isnan(log_prob)):
            raise ValueError("The initial log_prob was NaN")

        # emcee's RedBlueMove / StretchMove(a): shuffle the red/blue labels, then per half draw z, a partner from
        # the other half, and the acceptance threshold.  Written with index arrays, in-place arithmetic and masked
        # copies instead of boolean-mask indexing: on the small configurations the NumPy calls of this loop, not
        # the likelihood launches, set the pace (profiles/r02_configs.md).  The partner index is
        # floor(u * n_other) of one uniform draw rather than RandomState.randint (4 us instead of 9 per half).
        rng = self._random
        a_minus_1, inv_a, dim_minus_1 = self.a - 1.0, 1.0 / self.a, self.ndim - 1.0
        parity = np.arange(self.nwalkers) % 2
        naccepted = self.naccepted
        compute, uniform = self.compute_log_prob, rng.random_sample
        log, nonzero, take, copyto, intp = np.log, np.flatnonzero, np.take, np.copyto, np.intp
        with np.errstate(invalid='ignore', divide='ignore'):
            for _ in range(int(nsteps)):
                inds = parity.copy()
                rng.shuffle(inds)
                halves = (nonzero(inds == 0), nonzero(inds))
                for split in (0, 1):
                    i_s, i_c = halves[split], halves[1 - split]
                    ns = i_s.size
                    s = take(coords, i_s, axis=0)
                    zz = uniform(ns)
                    zz *= a_minus_1
                    zz += 1.0
                    zz *= zz
                    zz *= inv_a
                    pick = uniform(ns)
                    pick *= i_c.size
                    partner = take(coords, take(i_c, pick.astype(intp)), axis=0)
                    q = partner - s
                    q *= zz[:, None]
                    np.subtract(partner, q, out=q)                     # c_j - (c_j - s) z
                    new_log_prob = compute(q)
                    old_log_prob = take(log_prob, i_s)
                    lnpdiff = log(zz)
                    lnpdiff *= dim_minus_1
                    lnpdiff += new_log_prob
                    lnpdiff -= old_log_prob
                    accepted = lnpdiff > log(uniform(ns))
                    copyto(s, q, where=accepted[:, None])
                    copyto(old_log_prob, new_log_prob, where=accepted)
                    coords[i_s] = s
                    log_prob[i_s] = old_log_prob
                    naccepted[i_s] += accepted
                self.iteration += 1
                if store:
                    self._chain.append(coords.copy())
                    self._lnprob.append(log_prob.copy())
        return coords, log_prob, rng.get_state()


def make_host_sampler(nwalkers, ndim, log_prob_fn, seed=None):
    """emcee in vectorised mode when it is installed, else :class:`HostEnsembleSampler`."""
    try:
        import emcee
        emcee.moves.StretchMove          # a real emcee 3.x, not an import-only stub
    except (ImportError, AttributeError):
        return HostEnsembleSampler(nwalkers, ndim, log_prob_fn, seed=seed)
    sampler = emcee.EnsembleSampler(nwalkers, ndim, log_prob_fn, vectorize=True)
    if seed is not None:
        sampler._random = np.random.RandomState(seed)
    return sampler


class DeviceEnsembleSampler(object):
    """Stretch-move ensemble whose whole state lives on the GPU (``mcd_ensemble_*`` in
    ``include/mcd_b200.h``).  ``packed`` is the :class:`~mcmc_dynamics_b200.pack.PackedModel` whose
    ``lnprob`` the walkers sample."""

    def __init__(self, nwalkers, ndim, packed, a=2.0, seed=None):
        if nwalkers < 2 * ndim:
            raise RuntimeError("It is unadvisable to use a red-blue move with fewer walkers than twice the number "
                               "of dimensions.")
        if ndim != packed.n_theta:
            raise ValueError('ndim does not match the packed model')
        self.nwalkers = int(nwalkers)
        self.ndim = int(ndim)
        #: independent ensembles advanced together (one per segment of a segmented handle); state
        #: arrays then carry a leading segment axis
        self.nsegments = int(getattr(packed, 'n_segments', 1))
        self._packed = packed          # keeps the handle alive
        self._lib = _native.load_library()
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
        handle = ctypes.c_void_p()
        rc = self._lib.mcd_ensemble_create(packed.handle, self.nwalkers, ctypes.c_uint64(int(seed)), float(a),
                                           ctypes.byref(handle))
        if rc != 0:
            _native.check(rc)
        self._handle = handle
        self.iteration = 0
        self._chain = []
        self._lnprob = []
        self.naccepted = np.zeros(self._rows_shape, dtype=np.int64)
        self._have_state = False

    @property
    def _rows_shape(self):
        return (self.nsegments, self.nwalkers) if self.nsegments > 1 else (self.nwalkers,)

    @property
    def chain(self):
        """[nwalkers, nsteps, ndim] like emcee's; [nsegments, nwalkers, nsteps, ndim] when segmented."""
        if not self._chain:
            return np.empty(self._rows_shape + (0, self.ndim))
        steps_first = np.concatenate(self._chain, axis=0)              # [steps, (S,) W, P]
        return np.moveaxis(steps_first, 0, -2)

    @property
    def lnprobability(self):
        if not self._lnprob:
            return np.empty(self._rows_shape + (0,))
        return np.moveaxis(np.concatenate(self._lnprob, axis=0), 0, -1)

    def get_chain(self, discard=0, flat=False):
        chain = np.concatenate(self._chain, axis=0)[discard:] if self._chain else np.empty(
            (0,) + self._rows_shape + (self.ndim,))
        return chain.reshape((-1, self.ndim)) if flat else chain

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(max(1, self.iteration))

    def run_mcmc(self, initial_state, nsteps, log_prob0=None, rstate0=None, progress=False, store=True, **kwargs):
        """emcee's ``run_mcmc`` as far as ``analysis/runner.py:416-419`` uses it.  ``log_prob0`` is
        accepted and not used: the initial log-probabilities are always evaluated on the device (one
        launch), so a stale value handed in by the caller cannot enter the chain.
        ``rstate0`` must be ``None``: the random numbers are counter-based (Philox, keyed by the ``seed``
        given at construction and the step number), there is no generator state to restore; the state
        returned is ``None`` accordingly."""
        nsteps = int(nsteps)
        if rstate0 is not None:
            raise ValueError("DeviceEnsembleSampler draws counter-based random numbers keyed by its `seed`; "
                             "rstate0 cannot be applied")
        if initial_state is None and not self._have_state:
            raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
        if initial_state is not None and (not self._have_state or not np.array_equal(initial_state, self._last_pos)):
            pos = _native.contiguous(initial_state)
            if pos.shape != self._rows_shape + (self.ndim,):
                raise ValueError("incompatible input dimensions {0}".format(pos.shape))
            if not np.isfinite(pos).all():
                raise ValueError("At least one parameter value was infinite" if np.any(np.isinf(pos))
                                 else "At least one parameter value was NaN")
            rc = self._lib.mcd_ensemble_set_state(self._handle, _native.as_double_ptr(pos))
            if rc != 0:
                _native.check(rc)
            first = np.empty(self._rows_shape, dtype=np.float64)
            rc = self._lib.mcd_ensemble_get_state(self._handle, None, _native.as_double_ptr(first))
            if rc != 0:
                _native.check(rc)
            if np.any(np.isnan(first)):                   # emcee: checked before the first step, not after the last
                raise ValueError("The initial log_prob was NaN")
            self._have_state = True
        chain = np.empty((nsteps,) + self._rows_shape + (self.ndim,), dtype=np.float64) if store else None
        lnp = np.empty((nsteps,) + self._rows_shape, dtype=np.float64) if store else None
        nacc = np.zeros(self._rows_shape, dtype=np.int64)
        rc = self._lib.mcd_ensemble_run(
            self._handle, nsteps, _native.as_double_ptr(chain) if store else None,
            _native.as_double_ptr(lnp) if store else None, nacc.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        if rc != 0:
            _native.check(rc)
        self.naccepted = nacc
        self.iteration += nsteps
        if store:
            self._chain.append(chain)
            self._lnprob.append(lnp)
        pos = np.empty(self._rows_shape + (self.ndim,), dtype=np.float64)
        last = np.empty(self._rows_shape, dtype=np.float64)
        rc = self._lib.mcd_ensemble_get_state(self._handle, _native.as_double_ptr(pos), _native.as_double_ptr(last))
        if rc != 0:
            _native.check(rc)
        if np.any(np.isnan(last)):
            raise ValueError("Probability function returned NaN")
        self._last_pos = pos.copy()
        return pos, last, None

    @property
    def engine(self):
        """``(name, ctas_per_segment)`` of the last ``run_mcmc``: ``'resident'`` (whole run inside one
        kernel, catalogue in shared memory) or ``'graph'`` (CUDA graph of likelihood launches)."""
        kind, group = ctypes.c_int32(0), ctypes.c_int32(0)
        self._lib.mcd_ensemble_engine(self._handle, ctypes.byref(kind), ctypes.byref(group))
        return {0: None, 1: 'resident', 2: 'graph'}[kind.value], group.value

    def close(self):
        if self._handle is not None:
            self._lib.mcd_ensemble_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
