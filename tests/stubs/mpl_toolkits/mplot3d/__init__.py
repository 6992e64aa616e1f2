class Axes3D:  # noqa: D101 - import target only (toolplot.py:7,13)
    pass
