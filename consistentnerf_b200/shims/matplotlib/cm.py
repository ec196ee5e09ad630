"""cm stand-in: get_cmap(name) -> callable mapping [0,1] values to RGBA (a blue-green-red ramp)."""
import numpy as np


def get_cmap(_name=None, *_a, **_k):
    def cmap(x):
        x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
        r = np.clip(1.5 - np.abs(4 * x - 3), 0, 1)
        g = np.clip(1.5 - np.abs(4 * x - 2), 0, 1)
        b = np.clip(1.5 - np.abs(4 * x - 1), 0, 1)
        return np.stack([r, g, b, np.ones_like(x)], -1)
    return cmap
