"""Shim for the reference's `from loss import LossCalculator` (train.py:15)."""
from real_time_helmet_detection_b200.loss import LossCalculator  # noqa: F401
