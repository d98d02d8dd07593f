"""Drop-in for the reference's src/utils/transform_helpers.py: same builders and runners,
equal-shape measures batched into single launches."""
from wavelet_transformer_b200.api.transform_helpers import (create_cwt_dict, create_cwt_results_dict,  # noqa: F401
                                                            create_dwt_dict, create_dwt_regression_dict,
                                                            create_dwt_results_dict, create_xwt_dict,
                                                            create_xwt_results_dict)
