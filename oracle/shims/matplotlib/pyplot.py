"""TEST INFRASTRUCTURE ONLY: plotting is out of scope; attribute access raises AttributeError
so that introspection (doctest collection) keeps working."""


def __getattr__(name):
    raise AttributeError("matplotlib.pyplot.%s is stubbed in the oracle shims (plotting out of scope)" % name)
