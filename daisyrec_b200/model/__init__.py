from .MFRecommender import MF  # noqa: F401
