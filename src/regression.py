"""Drop-in for the numeric part of the reference's src/regression.py (plots and data retrieval
are out of scope): same function names, B200 engine underneath."""
from wavelet_transformer_b200.api.regression import *  # noqa: F401,F403
from wavelet_transformer_b200.api import regression as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
