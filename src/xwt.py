"""Drop-in for the reference's src/xwt.py: same public names, B200 engine underneath."""
from wavelet_transformer_b200.api.xwt import *  # noqa: F401,F403
from wavelet_transformer_b200.api import xwt as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
