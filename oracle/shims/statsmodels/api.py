from .regression.linear_model import OLS  # noqa: F401


def add_constant(data, prepend=True, has_constant='skip'):
    import numpy as np
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    ones = np.ones((data.shape[0], 1))
    return np.hstack([ones, data] if prepend else [data, ones])
