"""pyplot stand-in: every call is accepted and does nothing."""


def __getattr__(name):
    return lambda *a, **k: None
