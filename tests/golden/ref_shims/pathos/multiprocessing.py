class Pool(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('pathos stand-in: only importable')
