"""INTEGRATION.md Option B executed on the GPU, and SURVEY.md 8c KAT 9.

* The binding a maintainer would add to the reference (``tools/reference_binding.py``: ctypes over the C ABI +
  the patched ``Runner.lnprob``) is installed on objects with the reference Runner's attribute surface
  (``tests/binding_util.py``; ``tests/test_binding_cpu.py`` proves they compile the same descriptor as the
  reference's real objects) and must return what the UNPATCHED reference returned for the same inputs -- the
  golden vectors -- for all five classes, for one parameter vector and for an ``[n_walkers, n_free]`` array.
* KAT 9: the same host stretch move driven once by the NumPy oracle's ``lnprob`` (the reference's arithmetic)
  and once by the GPU ``lnprob`` from the same seed, on the mock of ``bin/run_tests.py:36-41``: posterior
  16/50/84 percentiles (``analysis/runner.py:566-613``) agree.
"""
import numpy as np
import pytest

import binding_util
import golden_util
from mcmc_dynamics_b200 import sampler as samplers
from mcmc_dynamics_b200 import synthetic
from mcmc_dynamics_b200.analysis import ModelFit
from oracle import harness

pytestmark = pytest.mark.gpu
GOLDEN = golden_util.load()
CASES = GOLDEN['cases']


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_patched_runner_lnprob_equals_the_unpatched_reference(case):
    rb = binding_util.binding_module()
    obj = binding_util.stand_in_for_case(case, GOLDEN)
    original = rb.install(type(obj))
    assert original is not rb.patched_lnprob
    theta = np.asarray(case['theta'])
    want = np.asarray(case['expected']['lnprob'])
    batch = obj.lnprob(theta)                                  # vectorised protocol: [W, P] -> [W]
    assert isinstance(batch, np.ndarray) and batch.shape == want.shape
    for k in range(len(theta)):
        single = obj.lnprob(theta[k])                          # the reference's protocol: one vector -> float
        assert isinstance(single, float)
        if np.isfinite(want[k]):
            assert single == pytest.approx(want[k], rel=1e-9, abs=0)
            assert batch[k] == pytest.approx(want[k], rel=1e-9, abs=0)
        else:
            assert single == want[k] == batch[k] == -np.inf
    # a parameter edited after the first call re-packs (bin/run_tests.py:88-93 edits between runs)
    free = obj.fitted_parameters
    name = 'sigma_max'
    obj.parameters[name].max = float(theta[0, free.index(name)]) - 1e-3
    assert obj.lnprob(theta[0]) == -np.inf
    rb.library().mcd_destroy(obj.__dict__['_b200'][1])


def test_kat9_oracle_driven_and_gpu_driven_chains_agree():
    data, truth = synthetic.mock_cluster(1500, seed=17)       # the mock of bin/run_tests.py:36-70
    model = ModelFit(data)
    model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    oracle = harness.oracle_for(model)
    n_walkers, n_steps, n_burn = 32, 260, 100
    pos = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=4, scale=0.1)
    gpu = samplers.HostEnsembleSampler(n_walkers, model.n_fitted_parameters, model.lnprob, seed=99)
    gpu.run_mcmc(pos, n_steps)
    cpu = samplers.HostEnsembleSampler(n_walkers, model.n_fitted_parameters, oracle.lnprob_many, seed=99)
    cpu.run_mcmc(pos, n_steps)
    pg = model.compute_percentiles(gpu.chain, n_burn)          # analysis/runner.py:566-613
    pc = model.compute_percentiles(cpu.chain, n_burn)
    # same seed and log-probabilities equal to ~1e-14: the two chains take the same decisions (a flip
    # needs |lnp difference| below the rounding difference), so the percentiles agree far inside the
    # Monte-Carlo error; the bound below is 1 % of the posterior width
    width = pg[2] - pg[0]
    assert np.all(np.abs(pg - pc) <= 0.01 * width), (pg, pc)
    assert np.allclose(gpu.lnprobability, cpu.lnprobability, rtol=1e-9, atol=0) or \
        np.mean(np.isclose(gpu.lnprobability, cpu.lnprobability, rtol=1e-9, atol=0)) > 0.9
    # and the posterior brackets the truth (mock recovery, the only known-answer scenario of the reference)
    for j, name in enumerate(model.fitted_parameters):
        assert abs(pg[1, j] - truth[name]) < 4.0 * max(width[j] / 2.0, 1e-9), (name, pg[:, j], truth[name])
