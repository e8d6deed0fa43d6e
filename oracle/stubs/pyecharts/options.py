def __getattr__(name):
    class _Opt:
        def __init__(self, *a, **k):
            pass
    _Opt.__name__ = name
    return _Opt
