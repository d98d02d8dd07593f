"""Drop-in for the reference's src/dwt.py: same public names, B200 engine underneath."""
from wavelet_transformer_b200.api.dwt import *  # noqa: F401,F403
from wavelet_transformer_b200.api import dwt as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
