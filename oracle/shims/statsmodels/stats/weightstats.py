"""Placeholder: associaTR imports this module at import time only (associaTR.py:15)."""
