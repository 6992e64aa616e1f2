rcParams = {}


def _noop(*a, **k):
    raise RuntimeError("matplotlib stub: plotting is not available in tests")


show = savefig = subplots = figure = plot = hist = _noop
