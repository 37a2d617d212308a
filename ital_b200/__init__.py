"""ital_b200: B200-native batch selection for Information-Theoretic Active Learning (drop-in `ITAL` learner)."""
from .learner import ITAL, EntropySampling, VarianceSampling  # noqa: F401

__all__ = ['ITAL', 'EntropySampling', 'VarianceSampling']
