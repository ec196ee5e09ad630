"""matplotlib stand-in: the reference imports pyplot / cm for figures it only draws in commented-out or test-set code."""
__version__ = "0-cnerf-shim"


def use(*_a, **_k):
    return None
