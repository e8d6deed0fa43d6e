class _Chart:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        def _f(*a, **k):
            return self
        return _f


class Surface3D(_Chart):
    pass


class Line3D(_Chart):
    pass


class Scatter3D(_Chart):
    pass


class Grid(_Chart):
    pass
