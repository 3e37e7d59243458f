"""Mirror of ``flux.config`` (reference src/flux/config.py:1-2)."""
DEBUG = False
DEFAULT_EPS = 1e-5
