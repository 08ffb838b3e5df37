"""Stand-in for lmfit.jsonutils: numbers, strings and None pass through; callables are described by
name like lmfit does -- objects without ``__name__`` (scipy.stats distribution instances) raise
AttributeError there too, which the reference's ``Parameters.dumps`` catches and skips."""


def encode4js(obj):
    if callable(obj):
        return dict(__class__='Callable', __name__=obj.__name__, importer=getattr(obj, '__module__', None))
    return obj


def decode4js(obj):
    if isinstance(obj, dict) and obj.get('__class__') == 'Callable':
        import importlib
        return getattr(importlib.import_module(obj['importer']), obj['__name__'])
    return obj
