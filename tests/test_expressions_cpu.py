"""The restricted evaluator behind ``initials`` / ``lnprior`` / ``expr`` strings (asteval in the reference,
``parameter.py:19-21,64-74,143,648,698``): same values as Python's own ``eval`` on the supported subset, and a
refusal -- not an evaluation -- for everything outside it."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from mcmc_dynamics_b200 import expressions

NAMES = {'a': 1.75, 'b': -0.5, 'c': 3.0, 'n': 4}


def _trees(depth):
    leaf = st.one_of(st.sampled_from(sorted(NAMES)), st.integers(-3, 9).map(str),
                     st.floats(0.25, 8.0, allow_nan=False).map(lambda x: repr(round(x, 3))))
    if depth == 0:
        return leaf
    sub = _trees(depth - 1)
    binary = st.tuples(sub, st.sampled_from(['+', '-', '*', '/']), sub).map(lambda t: '(%s %s %s)' % t)
    unary = sub.map(lambda s: '(-%s)' % s)
    compare = st.tuples(sub, st.sampled_from(['<', '<=', '>', '>=', '==', '!=']), sub).map(lambda t: '(%s %s %s)' % t)
    ternary = st.tuples(sub, compare, sub).map(lambda t: '(%s if %s else %s)' % t)
    call = st.tuples(st.sampled_from(['abs', 'sqrt', 'exp', 'cos', 'minimum', 'maximum']), sub, sub).map(
        lambda t: '%s(abs(%s) * 0.01 + 0.5%s)' % (t[0], t[1], ', ' + t[2] if t[0] in ('minimum', 'maximum') else ''))
    return st.one_of(leaf, binary, unary, ternary, call)


@settings(max_examples=300, deadline=None)
@given(_trees(3))
def test_same_value_as_python_eval(expression):
    symbols = dict(expressions.default_symbols(), **NAMES)
    try:
        with np.errstate(all='ignore'):
            want = eval(expression, {'__builtins__': {}}, symbols)     # noqa: S307 -- generated arithmetic only
    except ZeroDivisionError:
        with pytest.raises(ZeroDivisionError):
            expressions.evaluate(expressions.parse(expression), symbols)
        return
    with np.errstate(all='ignore'):
        got = expressions.evaluate(expressions.parse(expression), symbols)
    assert type(got) is type(want)
    assert (got == want) or (isinstance(want, float) and math.isnan(want) and math.isnan(got))


def test_names_are_the_dependencies():
    tree = expressions.parse('where(sigma_max > 2 * v_back, norm(0, 1).logpdf(val), -inf)')
    assert expressions.names(tree) == ['inf', 'norm', 'sigma_max', 'v_back', 'val', 'where']


@pytest.mark.parametrize('expression', [
    'lambda: 1', '[x for x in (1, 2)]', '(x for x in (1, 2))', '{x: 1 for x in (1, 2)}', 'a.__class__',
    '(a := 3)', 'a._private', 'import os', 'a = 3', '1 +',
])
def test_unsupported_syntax_is_refused_at_parse_time(expression):
    with pytest.raises(expressions.ExpressionError):
        expressions.parse(expression)


@pytest.mark.parametrize('expression', ['__import__("os")', 'open("/etc/passwd")', 'undefined_name + 1', 'eval("1")'])
def test_names_outside_the_symbol_table_do_not_resolve(expression):
    with pytest.raises(expressions.ExpressionError, match='is not defined'):
        expressions.evaluate(expressions.parse(expression), dict(expressions.default_symbols(), **NAMES))


def test_calls_attributes_indexing_and_slices():
    symbols = dict(expressions.default_symbols(), v=np.arange(6.0), rng=np.random.default_rng(3), n=5)
    assert expressions.evaluate(expressions.parse('v[1:4].sum() + v[-1]'), symbols) == 11.0
    assert expressions.evaluate(expressions.parse('maximum(v, 2.5)[0]'), symbols) == 2.5
    draws = expressions.evaluate(expressions.parse('rng.normal(loc=10, scale=0.1, size=n)'), symbols)
    assert draws.shape == (5,) and abs(draws.mean() - 10) < 0.3
    assert expressions.evaluate(expressions.parse('1 < 2 <= 2 and not (3 > 4) or False'), symbols) is True
