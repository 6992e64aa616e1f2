"""TEST-ONLY stub: the reference scripts import mpl_toolkits at module top; plotting is outside the hot path."""
