"""Drop-in for the reference's src/wct.py: same public names, B200 engine underneath."""
from wavelet_transformer_b200.api.wct import *  # noqa: F401,F403
from wavelet_transformer_b200.api import wct as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
