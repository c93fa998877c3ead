"""dsf_b200 - B200-native (sm_100a) implementation of DSF's differentiable model-fitting hot path."""
from .synthetic import make_synthetic_mano, sample_fit_inputs, write_mano_pkl  # noqa: F401

__all__ = ["make_synthetic_mano", "sample_fit_inputs", "write_mano_pkl"]
