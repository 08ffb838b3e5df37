"""CPU oracle for the mcmc-dynamics hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` may be imported by the product package
(``mcmc_dynamics_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only as
the checker / the timed CPU baseline.

How the oracle is pinned.  The reference ships no tests, golden vectors or
fixtures for this path, and it cannot be imported in this image (astropy, emcee,
asteval, lmfit, pathos, corner, matplotlib are absent; no network).  Parity is
therefore anchored on outputs of the reference's OWN source files:
``tests/golden/make_golden.py`` loads ``parameter.py``, ``analysis/runner.py``,
``constant.py``, ``model.py``, ``background/*.py``, ``calc_xy_offset.py`` and
``data_reader.py`` unmodified from ``/root/reference`` against minimal stand-ins
for the missing third-party modules (``tests/golden/ref_shims``) and records
``lnprior / lnlike / lnprob`` for all five model classes (fixed and free centre,
every background mode, prior rejections, ``no_sum``).  ``tests/test_golden_cpu.py``
checks this restatement against those vectors to 1e-12 relative.  Caveat, stated
here and in DESIGN.md: the stand-in for ``astropy.units`` restates astropy's
conversion rules; real astropy was never run, so the
unit handling is pinned only as far as that restatement is right.  Independent
hand-derived known answers: ``tests/test_oracle_kat.py``.
"""
