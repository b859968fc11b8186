"""TEST INFRASTRUCTURE ONLY: minimal stand-in for statsmodels (pinned 0.14.x in the
reference's pyproject.toml:82-87; absent from this image).  Only ``OLS(...).fit()`` is
restated (oracle/shims/statsmodels/regression/linear_model.py)."""
