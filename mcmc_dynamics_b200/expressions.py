"""Restricted evaluator for the ``initials`` / ``lnprior`` / ``expr`` strings of a Parameter.

The reference evaluates those strings with ``asteval`` against a per-``Parameters``
symbol table holding ``rng`` (a numpy Generator), ``rng_seed``, ``n``, ``val``,
``uniform``/``norm``/``lognorm`` from scipy.stats and the current value of every
parameter (``parameter.py:19-21,64-74,143,648,698``).  asteval is not available
here, so the same subset is interpreted directly from the Python AST: literals,
names from the symbol table, arithmetic, comparisons, conditional expressions,
attribute access and calls on symbol-table objects, tuples/lists and indexing.
Statements, imports, lambdas, comprehensions and dunder attributes are rejected.
"""
import ast
import math
import operator

import numpy as np

_BINOPS = {
    ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
    ast.FloorDiv: operator.floordiv, ast.Mod: operator.mod, ast.Pow: operator.pow,
}
_UNARYOPS = {ast.UAdd: operator.pos, ast.USub: operator.neg, ast.Not: operator.not_}
_CMPOPS = {
    ast.Eq: operator.eq, ast.NotEq: operator.ne, ast.Lt: operator.lt, ast.LtE: operator.le,
    ast.Gt: operator.gt, ast.GtE: operator.ge,
}

#: numpy names asteval exposes by default and that are plausible in a prior / initials string
_NUMPY_NAMES = ('pi', 'e', 'inf', 'nan', 'sqrt', 'exp', 'log', 'log10', 'log2', 'sin', 'cos', 'tan', 'arcsin',
                'arccos', 'arctan', 'arctan2', 'sinh', 'cosh', 'tanh', 'abs', 'fabs', 'floor', 'ceil', 'where',
                'minimum', 'maximum', 'isfinite', 'ones', 'zeros', 'full', 'linspace', 'arange', 'array',
                'deg2rad', 'rad2deg', 'hypot', 'sign', 'square', 'power', 'clip')


def default_symbols():
    table = {name: getattr(np, name) for name in _NUMPY_NAMES}
    table.update({'min': min, 'max': max, 'float': float, 'int': int, 'len': len, 'True': True,
                  'False': False, 'None': None, 'asin': math.asin, 'acos': math.acos, 'atan': math.atan,
                  'atan2': math.atan2})
    return table


class ExpressionError(ValueError):
    pass


def parse(expression):
    """Parse to an AST once; evaluation happens later against the current symbol table."""
    try:
        tree = ast.parse(expression.strip(), mode='eval')
    except SyntaxError as exc:
        raise ExpressionError("Cannot parse expression '{0}': {1}".format(expression, exc))
    for node in ast.walk(tree):
        if isinstance(node, (ast.Lambda, ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp,
                             ast.Await, ast.Yield, ast.YieldFrom, ast.NamedExpr)):
            raise ExpressionError("Unsupported syntax in expression '{0}'".format(expression))
        if isinstance(node, ast.Attribute) and node.attr.startswith('_'):
            raise ExpressionError("Private attribute access in expression '{0}'".format(expression))
    return tree


def names(tree):
    """Names an expression depends on (asteval's ``get_ast_names``)."""
    return sorted({node.id for node in ast.walk(tree) if isinstance(node, ast.Name)})


def evaluate(tree, symbols):
    return _eval(tree.body if isinstance(tree, ast.Expression) else tree, symbols)


def _eval(node, sym):
    if isinstance(node, ast.Constant):
        return node.value
    if isinstance(node, ast.Name):
        try:
            return sym[node.id]
        except KeyError:
            raise ExpressionError("name '{0}' is not defined".format(node.id))
    if isinstance(node, ast.BinOp):
        return _BINOPS[type(node.op)](_eval(node.left, sym), _eval(node.right, sym))
    if isinstance(node, ast.UnaryOp):
        return _UNARYOPS[type(node.op)](_eval(node.operand, sym))
    if isinstance(node, ast.BoolOp):
        values = [_eval(v, sym) for v in node.values]
        if isinstance(node.op, ast.And):
            out = values[0]
            for v in values[1:]:
                out = out and v
            return out
        out = values[0]
        for v in values[1:]:
            out = out or v
        return out
    if isinstance(node, ast.Compare):
        left = _eval(node.left, sym)
        result = True
        for op, comparator in zip(node.ops, node.comparators):
            right = _eval(comparator, sym)
            result = result and _CMPOPS[type(op)](left, right)
            left = right
        return result
    if isinstance(node, ast.IfExp):
        return _eval(node.body, sym) if _eval(node.test, sym) else _eval(node.orelse, sym)
    if isinstance(node, ast.Attribute):
        return getattr(_eval(node.value, sym), node.attr)
    if isinstance(node, ast.Call):
        func = _eval(node.func, sym)
        args = [_eval(a, sym) for a in node.args]
        kwargs = {kw.arg: _eval(kw.value, sym) for kw in node.keywords}
        return func(*args, **kwargs)
    if isinstance(node, (ast.Tuple, ast.List)):
        return [_eval(e, sym) for e in node.elts]
    if isinstance(node, ast.Subscript):
        return _eval(node.value, sym)[_eval(node.slice, sym)]
    if isinstance(node, ast.Slice):
        return slice(None if node.lower is None else _eval(node.lower, sym),
                     None if node.upper is None else _eval(node.upper, sym),
                     None if node.step is None else _eval(node.step, sym))
    raise ExpressionError('Unsupported expression element: {0}'.format(type(node).__name__))
