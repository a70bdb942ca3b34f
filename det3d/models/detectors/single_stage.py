"""Drop-in for reference det3d/models/detectors/single_stage.py (implemented in pillarnext_b200.modules)."""
from pillarnext_b200.modules import SingleStageDetector  # noqa: F401
