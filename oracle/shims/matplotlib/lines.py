"""TEST INFRASTRUCTURE ONLY: import-time stub (compareSTR.py:20 imports Line2D for its legends)."""


class Line2D:
    def __init__(self, *a, **k):
        pass
