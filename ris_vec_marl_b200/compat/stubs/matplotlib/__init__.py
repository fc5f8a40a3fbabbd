"""No-op stand-in: the reference drivers import matplotlib only to save reward plots, and
matplotlib is not installed in this image."""
rcParams = {}


def use(*a, **k):
    pass
