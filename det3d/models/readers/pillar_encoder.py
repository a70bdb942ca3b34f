"""Drop-in for reference det3d/models/readers/pillar_encoder.py (implemented in pillarnext_b200.modules)."""
from pillarnext_b200.modules import PFNLayer, PillarFeatureNet, PillarNet  # noqa: F401
