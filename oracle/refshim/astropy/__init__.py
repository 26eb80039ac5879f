"""Test-only stand-in for the one thing the reference takes from astropy: CODATA constants (phys_const.py:24-44)."""
from . import constants  # noqa: F401
