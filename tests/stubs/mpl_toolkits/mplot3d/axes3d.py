from . import Axes3D  # noqa: F401
