"""
TEST INFRASTRUCTURE ONLY.  numpy restatement of the published algorithm of
``statsmodels.regression.linear_model.OLS(endog, exog, missing='drop').fit()``
(statsmodels 0.14, method='pinv'), the only statsmodels entry point on the hot
path (reference call site trtools/associaTR/associaTR.py:281-290):

* rows with any NaN in endog/exog are dropped (missing='drop');
* ``pinv_wexog = pinv(X)`` (SVD, rcond=1e-15); ``params = pinv_wexog @ y``;
* ``normalized_cov_params = pinv_wexog @ pinv_wexog.T``;
* ``rank = matrix_rank(diag(singular values))``; ``df_resid = n - rank``;
* ``scale = ssr / df_resid``; ``bse = sqrt(diag(cov * scale))``;
* ``tvalues = params / bse``; ``pvalues = 2 * t.sf(|t|, df_resid)``;
* ``rsquared = 1 - ssr/centered_tss`` when the design holds a constant column
  (detected as a column with zero peak-to-peak range and a non-zero value), else
  ``1 - ssr/uncentered_tss``.

Pinned against the independent plink2 --glm goldens the reference's own tests use
(trtools/associaTR/tests/test_associaTR.py:57-86) — see tests/test_oracle_golden.py.
"""
import numpy as np
import scipy.stats


class _Results:
    pass


class OLS:
    def __init__(self, endog, exog, missing='none', hasconst=None):
        y = np.asarray(endog, dtype=float).reshape(-1)
        X = np.asarray(exog, dtype=float)
        if X.ndim == 1:
            X = X[:, None]
        if missing == 'drop':
            keep = ~(np.isnan(y) | np.any(np.isnan(X), axis=1))
            y, X = y[keep], X[keep]
        self.endog, self.exog = y, X
        # statsmodels' constant detection (base/data.py _handle_constant)
        if X.shape[0] > 0:
            ptp = np.ptp(X, axis=0)
            const_cols = (ptp == 0) & np.all(X != 0, axis=0)
            self.k_constant = int(np.any(const_cols))
        else:
            self.k_constant = 0

    def fit(self):
        y, X = self.endog, self.exog
        res = _Results()
        pinv, sv = np.linalg.pinv(X, rcond=1e-15), np.linalg.svd(X, compute_uv=False)
        params = pinv @ y
        ncp = pinv @ pinv.T
        rank = np.linalg.matrix_rank(np.diag(sv))
        n = X.shape[0]
        df_resid = n - rank
        resid = y - X @ params
        ssr = float(resid @ resid)
        scale = ssr / df_resid
        bse = np.sqrt(np.diag(ncp * scale))
        tvalues = params / bse
        res.params = params
        res.bse = bse
        res.tvalues = tvalues
        res.pvalues = scipy.stats.t.sf(np.abs(tvalues), df_resid) * 2
        res.df_resid = df_resid
        res.ssr = ssr
        if self.k_constant:
            tss = float(np.sum((y - y.mean()) ** 2))
        else:
            tss = float(y @ y)
        res.rsquared = 1 - ssr / tss
        res.resid = resid
        return res
