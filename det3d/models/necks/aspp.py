"""Drop-in for reference det3d/models/necks/aspp.py (implemented in pillarnext_b200.modules)."""
from pillarnext_b200.modules import ASPPNeck  # noqa: F401
