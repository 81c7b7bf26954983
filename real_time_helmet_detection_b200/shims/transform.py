"""Shim for the reference's `from transform import hm2box` (evaluate.py:13, export.py) / `box2hm` (data.py:14)."""
from real_time_helmet_detection_b200.transform import hm2box, box2hm  # noqa: F401
