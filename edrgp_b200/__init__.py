"""edrgp_b200 -- B200-native (sm_100a) sparse-GP posterior-gradient EDR hot path.

Drop-in for edr-gp's ``SparseGaussianProcessRegressor`` + ``SVDTransformer`` +
``EffectiveDimensionalityReduction`` on that path; see DESIGN.md and INTEGRATION.md.
The CUDA library (``libedrgp_b200.so``, C ABI in ``include/edrgp_b200.h``) is loaded on first use;
there is no CPU fallback.
"""
__version__ = '0.1.0'

from .gp_model import SparseGaussianProcessRegressor          # noqa: F401
from .transformer import GramEighTransformer, DevicePCA       # noqa: F401
from .edr import EffectiveDimensionalityReduction, EDR, BlockEDR   # noqa: F401
from .utils import (discrepancy, ort_space, subspace_variance_ratio,   # noqa: F401
                    subspace_variance_ratio_from_gram)
from . import datasets                                        # noqa: F401

SVDTransformer = GramEighTransformer      # the reference's name for the transformer this one replaces (edrgp/utils.py:81)
