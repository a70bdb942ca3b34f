"""Drop-in for reference det3d/models/utils/conv.py."""
from pillarnext_b200.modules import BasicBlock, ConvBlock  # noqa: F401
