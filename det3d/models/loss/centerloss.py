"""Drop-in for the loss functions of reference det3d/models/loss/centerloss.py."""
from pillarnext_b200.loss import diou_aligned as bbox3d_overlaps_diou  # noqa: F401
from pillarnext_b200.loss import fast_focal_loss, iou_reg_loss, reg_loss  # noqa: F401
