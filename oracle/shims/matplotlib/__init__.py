"""TEST INFRASTRUCTURE ONLY: import-time stub (statSTR.py:6-13 sets a backend and rcParams)."""
rcParams = {}


def use(*a, **k):
    pass
