"""Device-resident ensemble sampler (csrc/mcd_sampler.cu) and the end-to-end sampling call."""
import numpy as np
import pytest

from common import build
from mcmc_dynamics_b200 import sampler as samplers
from mcmc_dynamics_b200 import synthetic
from mcmc_dynamics_b200.analysis import ConstantFit, ModelFit

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['resident', 'resident-1', 'resident-7', 'graph'])
def sampler_path(request, monkeypatch):
    """The device-sampler engines: whole chains inside one kernel with the catalogue in shared memory
    (group size chosen by the library, one CTA, seven CTAs with ragged star slices), and the CUDA
    graph of fused likelihood launches (any size; forced with MCD_NO_RESIDENT_CHAIN=1)."""
    monkeypatch.setenv('MCD_NO_RESIDENT_CHAIN', '1' if request.param == 'graph' else '0')
    monkeypatch.setenv('MCD_FORCE_RESIDENT_CHAIN', '1' if request.param.startswith('resident') else '0')
    if '-' in request.param:
        monkeypatch.setenv('MCD_CHAIN_GROUP', request.param.split('-')[1])
    else:
        monkeypatch.delenv('MCD_CHAIN_GROUP', raising=False)
    return request.param


def _mock_model(n_stars=1500, seed=21, cls=ModelFit):
    data, truth = synthetic.mock_cluster(n_stars, seed=seed)
    model = cls(data)
    model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    return model, truth


def test_device_chain_is_self_consistent_and_reproducible(sampler_path):
    model, truth = _mock_model()
    pos = synthetic.initial_ball(truth, model.fitted_parameters, 32, seed=3)
    runs = []
    for _ in range(2):
        s = samplers.DeviceEnsembleSampler(32, model.n_fitted_parameters, model.pack(), seed=1234)
        s.run_mcmc(pos, 40)
        runs.append((s.chain.copy(), s.lnprobability.copy(), s.naccepted.copy()))
        kind, group = s.engine
        assert kind == sampler_path.split('-')[0]
        if '-' in sampler_path:
            assert group == int(sampler_path.split('-')[1])
    chain, lnp, nacc = runs[0]
    assert chain.shape == (32, 40, 6) and lnp.shape == (32, 40)
    assert np.array_equal(chain, runs[1][0]) and np.array_equal(lnp, runs[1][1])     # same seed, same chain
    # the stored log-probabilities belong to the stored positions
    for step in (0, 17, 39):
        again = model.lnprob(np.ascontiguousarray(chain[:, step, :]))
        assert np.allclose(again, lnp[:, step], rtol=1e-12, atol=0)
    assert np.all(np.isfinite(lnp))
    frac = nacc / 40.0
    assert 0.1 < frac.mean() < 0.95
    # walkers only ever move to accepted proposals: consecutive equal positions <=> equal lnprob
    moved = np.any(chain[:, 1:, :] != chain[:, :-1, :], axis=2)
    assert np.array_equal(moved, lnp[:, 1:] != lnp[:, :-1])
    assert moved.sum() == nacc.sum() - np.any(chain[:, 0, :] != pos, axis=1).sum()


def test_device_and_host_samplers_agree_and_recover_the_truth(sampler_path):
    """Mock recovery (bin/run_tests.py:36-41 scenario): posterior percentiles (runner.py:566-613) of
    the device sampler and of the host stretch move driving the same GPU lnprob agree within
    Monte-Carlo error, and bracket the truth."""
    model, truth = _mock_model(n_stars=4000, seed=8)
    n_walkers, n_steps, n_burn = 64, 700, 250
    pos = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=4, scale=0.1)
    dev = model(n_walkers=n_walkers, n_steps=n_steps, pos=pos, sampler='device', seed=11, prefix=None)
    host = model(n_walkers=n_walkers, n_steps=n_steps, pos=pos, sampler='host', seed=12, prefix=None)
    assert dev.chain.shape == host.chain.shape == (n_walkers, n_steps, 6)
    pd = model.compute_percentiles(dev.chain, n_burn)        # [3, n_fitted] like runner.py:566-613
    ph = model.compute_percentiles(host.chain, n_burn)
    assert pd.shape == (3, 6)
    for j, name in enumerate(model.fitted_parameters):
        lo_d, med_d, hi_d = pd[:, j]
        lo_h, med_h, hi_h = ph[:, j]
        width = 0.5 * ((hi_d - lo_d) + (hi_h - lo_h)) / 2.0
        assert abs(med_d - med_h) < 0.35 * 2 * width, (name, pd[:, j], ph[:, j])
        assert 0.6 < (hi_d - lo_d) / (hi_h - lo_h) < 1.6, (name, pd[:, j], ph[:, j])
        # truth within ~3.5 sigma of the posterior
        assert abs(med_d - truth[name]) < 3.5 * width + 1e-9, (name, pd[:, j], truth[name])


def test_runner_call_signature_and_checkpoint(tmp_path):
    model, truth = _mock_model(n_stars=300, cls=ConstantFit)
    model.parameters['sigma_max'].set(initials='rng.lognormal(mean=2.3, sigma=0.3, size=n)')
    prefix = str(tmp_path / 'run')
    s = model(n_walkers=16, n_steps=30, n_out=10, prefix=prefix)
    assert s.iteration == 30 and s.chain.shape == (16, 30, 4)
    chain = model.read_chain(prefix + '_chain.pkl')
    assert chain.shape == (16, 30, 4)
    last = model.read_final_chain(prefix + '_chain.pkl')
    assert np.array_equal(last, s.chain[:, -1, :])
    # resume from the stored positions (run.py:419-422,450)
    s2 = model(n_walkers=16, n_steps=5, pos=last, prefix=None)
    assert s2.chain.shape == (16, 5, 4)
    bad = last.copy()
    bad[3, model.fitted_parameters.index('sigma_max')] = -1.0
    with pytest.raises(ValueError, match='Invalid initial guesses'):
        model(n_walkers=16, n_steps=5, pos=bad, prefix=None)


def test_emulated_star_shards_add_up():
    """Two shard handles on one GPU: partial sums add up to the whole-catalogue lnprob (the
    multi-GPU data path minus the collective)."""
    import torch
    columns, truth = synthetic.mock_cluster(5001, seed=5, as_reader=False)
    from mcmc_dynamics_b200 import sharded

    def make(cols):
        m = ModelFit(synthetic.reader_from_columns(cols))
        m.parameters['ra_center'].set(value=truth['ra_center'])
        m.parameters['dec_center'].set(value=truth['dec_center'])
        return m
    whole = make(columns)
    theta = synthetic.initial_ball(truth, whole.fitted_parameters, 40, seed=6)
    theta[7, whole.fitted_parameters.index('a')] = -1.0
    th = torch.as_tensor(theta, device='cuda:0')
    total = whole.lnprob_tensor(th)
    parts = sum(make(sharded.shard_columns(columns, r, 3)).pack().lnprob_partial_tensor(th) for r in range(3))
    assert parts[7] == -np.inf and total[7] == -np.inf
    keep = np.arange(40) != 7
    assert np.allclose(parts.cpu().numpy()[keep], total.cpu().numpy()[keep], rtol=1e-12, atol=0)


@pytest.mark.parametrize('n_walkers', [600, 1024])
def test_many_walkers_several_walker_groups(n_walkers, sampler_path):
    """More than 256 active walkers per half-step: several walker groups per launch, each accepting in
    place in the fused kernel; the resident kernel covers up to 512 active walkers per CTA."""
    model, truth = _mock_model(n_stars=2500, seed=31)
    pos = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=9, scale=0.1)
    s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=5)
    s.run_mcmc(pos, 12)
    chain, lnp = s.chain, s.lnprobability
    assert chain.shape == (n_walkers, 12, 6) and np.all(np.isfinite(lnp))
    for step in (0, 11):
        again = model.lnprob(np.ascontiguousarray(chain[:, step, :]))
        assert np.allclose(again, lnp[:, step], rtol=1e-12, atol=0)
    moved = np.any(chain[:, 1:, :] != chain[:, :-1, :], axis=2)
    assert np.array_equal(moved, lnp[:, 1:] != lnp[:, :-1])
    assert 0.1 < (s.naccepted / 12.0).mean() < 0.95


def test_engine_choice_falls_back_to_the_graph_of_launches(monkeypatch):
    """What the resident kernel cannot hold goes to the launch engine without the caller noticing: more
    than 1024 walkers (a half-ensemble must fit one CTA), or an odd ensemble on either engine."""
    monkeypatch.setenv('MCD_FORCE_RESIDENT_CHAIN', '1')
    monkeypatch.setenv('MCD_NO_RESIDENT_CHAIN', '0')
    monkeypatch.delenv('MCD_CHAIN_GROUP', raising=False)
    model, truth = _mock_model(n_stars=1200, seed=33)
    for n_walkers, engine in ((1100, 'graph'), (33, 'resident')):
        pos = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=2, scale=0.1)
        s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=8)
        s.run_mcmc(pos, 6)
        assert s.engine[0] == engine
        chain, lnp = s.chain, s.lnprobability
        again = model.lnprob(np.ascontiguousarray(chain[:, -1, :]))
        assert np.allclose(again, lnp[:, -1], rtol=1e-12, atol=0)
        assert s.naccepted.sum() > 0


@pytest.mark.parametrize('cls,n_stars,n_walkers', [(ConstantFit, 3001, 16), (ModelFit, 10_000, 128), (ModelFit, 40_000, 32)])
def test_resident_chain_groups_walk_the_same_chain(cls, n_stars, n_walkers, monkeypatch):
    """One CTA, a ragged group and a whole-GPU group hold different slices of the stars but draw the same
    proposals from the same counters: their chains agree to rounding of the slice sums until the
    first accept/reject decision that falls inside that rounding (none in these few steps)."""
    model, truth = _mock_model(n_stars=n_stars, seed=31, cls=cls)
    pos = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=6)
    monkeypatch.setenv('MCD_FORCE_RESIDENT_CHAIN', '1')
    monkeypatch.setenv('MCD_NO_RESIDENT_CHAIN', '0')
    chains = {}
    for group, exchange in ((1, 'tagged'), (7, 'tagged'), (7, 'counter'), (40, 'tagged'), (148, 'tagged'), (148, 'counter')):
        if n_stars / group > 6000:
            continue                      # a slice of that size does not fit one SM's shared memory
        monkeypatch.setenv('MCD_CHAIN_GROUP', str(group))
        monkeypatch.setenv('MCD_CHAIN_EXCHANGE', exchange)     # how the CTAs of a group trade their sums
        s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=99)
        s.run_mcmc(pos, 12)
        kind, used = s.engine
        if kind != 'resident':
            continue                      # fewer SMs than the group asks for
        assert used == group
        chains[group, exchange] = (s.chain.copy(), s.lnprobability.copy())
    assert len(chains) >= 2
    ref_chain, ref_lnp = chains[sorted(chains)[0]]
    for key, (chain, lnp) in chains.items():
        assert np.allclose(chain, ref_chain, rtol=1e-10, atol=0), key
        assert np.allclose(lnp, ref_lnp, rtol=1e-12, atol=0), key
    # the two ways of trading sums add them in the same order: identical chains
    for group in (7, 148):
        if (group, 'tagged') in chains and (group, 'counter') in chains:
            assert np.array_equal(chains[group, 'tagged'][0], chains[group, 'counter'][0])
    again = model.lnprob(np.ascontiguousarray(ref_chain[:, -1, :]))
    assert np.allclose(again, ref_lnp[:, -1], rtol=1e-12, atol=0)


@pytest.mark.parametrize('variant', ['ConstantFit+bg', 'ConstantFitGB', 'ModelFitGB', 'ModelFitConstantBackground'])
def test_device_sampler_on_the_background_variants(variant, sampler_path):
    """Every mixture kernel inside both sampler engines: the stored log-probabilities are the model's
    lnprob of the stored positions (checked against the oracle too), walkers move and stay inside the
    box prior."""
    from oracle import harness
    model, oracle, theta, truth = build(variant, n_stars=1500, seed=3)
    n_walkers = 4 * model.n_fitted_parameters
    pos = theta(n_walkers, seed=11, scale=0.05)
    s = samplers.DeviceEnsembleSampler(n_walkers, model.n_fitted_parameters, model.pack(), seed=21)
    s.run_mcmc(pos, 25)
    assert s.engine[0] == sampler_path.split('-')[0]
    chain, lnp = s.chain, s.lnprobability
    assert np.all(np.isfinite(lnp))
    last = np.ascontiguousarray(chain[:, -1, :])
    assert np.allclose(model.lnprob(last), lnp[:, -1], rtol=1e-12, atol=0)
    assert harness.relative_error(lnp[:8, -1], oracle.lnprob_many(last[:8])) < 1e-9
    assert 0.05 < (s.naccepted / 25.0).mean() < 0.95


def test_graph_sampler_survives_scratch_growth_and_repack(monkeypatch):
    """A captured ensemble graph bakes in the handle's scratch buffers, packed columns and routing.  A
    larger lnprob call (scratch re-allocated) or a re-pack between two ``run_mcmc`` calls must make the
    sampler re-capture instead of replaying through freed pointers: the continued chain equals an
    uninterrupted one bit for bit."""
    monkeypatch.setenv('MCD_NO_RESIDENT_CHAIN', '1')
    model, truth = _mock_model(n_stars=6000, seed=31)
    pos = synthetic.initial_ball(truth, model.fitted_parameters, 24, seed=3)
    ref = samplers.DeviceEnsembleSampler(24, model.n_fitted_parameters, model.pack(), seed=77)
    ref.run_mcmc(pos, 30)
    s = samplers.DeviceEnsembleSampler(24, model.n_fitted_parameters, model.pack(), seed=77)
    s.run_mcmc(pos, 10)
    assert s.engine[0] == 'graph'
    big = synthetic.initial_ball(truth, model.fitted_parameters, 2000, seed=9)
    lnp_big = model.lnprob(big)                        # many more walkers: partials / counters grow
    assert np.all(np.isfinite(lnp_big))
    s.run_mcmc(None, 10)
    model.math_mode = 'plain'                          # re-pack (new kernel variant, columns rewritten) ...
    model.lnprob(big[:3])
    model.math_mode = 'fast'                           # ... and back
    model.lnprob(big[:3])
    s.run_mcmc(None, 10)
    assert np.array_equal(s.chain, ref.chain) and np.array_equal(s.lnprobability, ref.lnprobability)


def test_device_sampler_rejects_bad_initial_state():
    """emcee raises before the first step when the initial log-probability is NaN or a coordinate is not
    finite; so does the device sampler, and it refuses an `rstate0`."""
    model, truth = _mock_model(n_stars=200)
    pos = synthetic.initial_ball(truth, model.fitted_parameters, 16, seed=3)
    s = samplers.DeviceEnsembleSampler(16, model.n_fitted_parameters, model.pack(), seed=1)
    bad = pos.copy()
    bad[2, 0] = np.nan
    with pytest.raises(ValueError, match='NaN'):
        s.run_mcmc(bad, 2)
    bad[2, 0] = np.inf
    with pytest.raises(ValueError, match='infinite'):
        s.run_mcmc(bad, 2)
    with pytest.raises(ValueError, match='rstate0'):
        s.run_mcmc(pos, 2, rstate0=np.random.RandomState(1).get_state())
    with pytest.raises(ValueError, match='initial_state'):
        s.run_mcmc(None, 2)
    out = s.run_mcmc(pos, 2, log_prob0=np.zeros(16))   # accepted, recomputed on the device
    assert np.allclose(model.lnprob(out[0]), out[1], rtol=1e-12)
