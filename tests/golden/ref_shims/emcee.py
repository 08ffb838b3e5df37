class EnsembleSampler(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('emcee stand-in: only importable')
