"""Headless no-op stand-in for matplotlib (the reference imports it at module import time and
calls plt.show() after a run, plots.py:50-51)."""
from . import pyplot  # noqa: F401


def use(*a, **k):
    pass
