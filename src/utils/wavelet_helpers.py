"""Drop-in for the numeric helpers of the reference's src/utils/wavelet_helpers.py."""
from wavelet_transformer_b200.api.wavelet_helpers import (align_series, normalize_xwt_results,  # noqa: F401
                                                          standardize_series)
