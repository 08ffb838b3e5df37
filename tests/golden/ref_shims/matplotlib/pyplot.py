class _Anything(object):
    def __getattr__(self, name):
        raise NotImplementedError('matplotlib stand-in: only importable')


def __getattr__(name):
    return _Anything
