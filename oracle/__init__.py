"""CPU oracle for the mcmc-dynamics hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` may be imported by the product package
(``mcmc_dynamics_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only as
the checker / the timed CPU baseline.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path, and it cannot be imported in this image (astropy, emcee, asteval,
lmfit are absent and there is no network).  The oracle is therefore pinned only
by (i) op-for-op restatement of the cited reference lines, (ii) hand-derived
known-answer tests in ``tests/test_oracle_kat.py`` and (iii) the reference's own
source files executed against minimal stand-ins for the missing third-party
modules (``tests/golden/make_golden.py``).
"""
