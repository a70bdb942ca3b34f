"""Drop-in for reference det3d/models/heads/centerhead.py (implemented in pillarnext_b200.modules)."""
from pillarnext_b200.modules import CenterHead, SepHead  # noqa: F401
