class DistributedDataParallelKwargs:
    def __init__(self, **kw):
        self.kw = kw
