"""INTEGRATION.md Option B, host half: the descriptor the binding (tools/reference_binding.py) compiles.

* from the stand-ins the GPU tests use == from the REAL reference objects (stored in the golden file when it
  was generated from /root/reference's own classes, tests/golden/make_golden.py),
* == what the product's own packing (mcmc_dynamics_b200/pack.py) compiles for the mirrored model,
* and, where /root/reference exists (this container, not the GPU box), re-derived live from the reference's
  classes.
No GPU: nothing here creates a device handle.
"""
import os
import sys

import numpy as np
import pytest

import binding_util
import golden_util

GOLDEN = golden_util.load()
CASES = GOLDEN['cases']
IDS = [c['name'] for c in CASES]


def same_descriptor(a, b):
    for key in ('rotation', 'background', 'n_theta', 'n_stars', 'slot', 'fixed_prior_ok'):
        assert a[key] == b[key], key
    for key in ('unit_scale', 'lower', 'upper'):
        assert np.array_equal(np.asarray(a[key], dtype=float), np.asarray(b[key], dtype=float)), key
    # fixed values matter for the fixed slots only (a sampled slot's entry is the free-centre expansion point)
    for k, slot in enumerate(a['slot']):
        if slot < 0:
            assert a['fixed_value'][k] == pytest.approx(b['fixed_value'][k], rel=1e-15, abs=0), k


@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_stand_in_compiles_the_descriptor_of_the_real_reference_object(case):
    rb = binding_util.binding_module()
    obj = binding_util.stand_in_for_case(case, GOLDEN)
    desc, keep = rb.describe(obj)
    same_descriptor(rb.summary(desc), case['binding_descriptor'])
    assert set(keep) >= {'ra', 'dec', 'v', 'verr'} and all(a.dtype == np.float64 for a in keep.values())


@pytest.mark.parametrize('case', CASES, ids=IDS)
def test_product_packing_equals_the_binding_descriptor(case):
    """The mirror classes route parameters, units and bounds exactly as the binding does on the reference."""
    from mcmc_dynamics_b200 import _native
    # the mirror computes its background column on the GPU: hand it the reference's column instead
    model = golden_util.product_for_case_host_only(case)
    desc, _ = model._descriptor()
    want = case['binding_descriptor']
    got = {'rotation': desc.rotation, 'background': desc.background, 'n_theta': desc.n_theta, 'n_stars': desc.n_stars,
           'slot': list(desc.slot), 'fixed_value': list(desc.fixed_value), 'unit_scale': list(desc.unit_scale),
           'lower': list(desc.lower)[:desc.n_theta], 'upper': list(desc.upper)[:desc.n_theta],
           'fixed_prior_ok': desc.fixed_prior_ok}
    same_descriptor(got, want)
    assert len(_native.PARAM_SLOTS) == len(want['slot'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/mcmc_dynamics'), reason='needs the reference sources')
def test_live_reference_objects_give_the_stored_descriptors():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden
    R = make_golden.load_reference()
    rb = binding_util.binding_module()
    u = R.u
    case = next(c for c in CASES if c['name'] == 'ModelFit v_sys fixed, bounded')
    cols = golden_util.case_columns(case)
    units = {'ra': u.deg, 'dec': u.deg, 'v': u.km / u.s, 'verr': u.km / u.s}
    data = R.DataReader({k: u.Quantity(v, units[k]) for k, v in cols.items()})
    model = R.ModelFit(data)
    for name, edit in case['parameter_edits'].items():
        edit = dict(edit)
        if 'value' in edit:
            edit['value'] = u.Quantity(edit['value'], model.parameters[name].unit)
        model.parameters[name].set(**edit)
    same_descriptor(rb.summary(rb.describe(model)[0]), case['binding_descriptor'])
    # install() swaps the method on the reference's own Runner class
    from mcmc_dynamics.analysis.runner import Runner
    original = rb.install(Runner)
    try:
        assert Runner.lnprob is rb.patched_lnprob and R.ModelFit.lnprob is rb.patched_lnprob
    finally:
        Runner.lnprob = original
