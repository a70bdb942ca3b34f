"""Drop-in for the 2-D classes of reference det3d/models/utils/sparse_conv.py."""
from pillarnext_b200.modules import SparseBasicBlock, SparseConvBlock  # noqa: F401
