"""The plain-C restatement (oracle/oracle_c.c) against the golden vectors of the reference's own
source files and against the NumPy restatement: two independently written checkers must agree."""
import numpy as np
import pytest

import golden_util
from oracle import c_port
from oracle import reference_np as ref

pytestmark = pytest.mark.skipif(not c_port.available(), reason='oracle/_build/liboracle_c.so not built')
GOLDEN = golden_util.load()


@pytest.mark.parametrize('case', GOLDEN['cases'], ids=[c['name'] for c in GOLDEN['cases']])
def test_c_oracle_reproduces_reference_outputs(case):
    oracle = golden_util.oracle_for_case(case)
    c = c_port.COracle(oracle)
    theta = np.asarray(case['theta'])
    got = c.lnprob_many(theta)
    want = np.asarray(case['expected']['lnprob'])
    fin = np.isfinite(want)
    assert np.array_equal(np.isinf(got), ~fin)
    assert np.allclose(got[fin], want[fin], rtol=1e-11, atol=0)
    with np.errstate(all='ignore'):
        assert np.allclose(got[fin], oracle.lnprob_many(theta)[fin], rtol=1e-11, atol=0)


def test_c_backgrounds_match_numpy():
    rng = np.random.default_rng(3)
    v, verr, v_bg = rng.normal(0, 50, 300), rng.uniform(0.5, 6, 300), rng.normal(10, 40, 77)
    assert np.allclose(c_port.gaussian_background(v, verr, 3.0, 20.0), ref.gaussian_background(v, verr, 3.0, 20.0),
                       rtol=1e-13)
    assert np.allclose(c_port.single_stars_background(v_bg, v, verr, 2.0),
                       ref.single_stars_background(v_bg, v, verr, 2.0), rtol=1e-12)
