"""Drop-in for reference det3d/models/backbones/sparse_resnet.py (implemented in pillarnext_b200.modules)."""
from pillarnext_b200.modules import SparseResNet  # noqa: F401
