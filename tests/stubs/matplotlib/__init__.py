"""Test-only stub: the unmodified reference scripts `import matplotlib.pyplot` at module top (driver.py:8,
monte_carlo.py:7, sensitivity.py:8) and matplotlib is not installed here.  Nothing is plotted in tests."""
rcParams = {}
