"""edrgp_b200 -- B200-native (sm_100a) sparse-GP posterior-gradient EDR hot path.

Drop-in for edr-gp's ``SparseGaussianProcessRegressor`` + ``SVDTransformer`` +
``EffectiveDimensionalityReduction`` on that path; see DESIGN.md and INTEGRATION.md.
"""
__version__ = '0.1.0'
