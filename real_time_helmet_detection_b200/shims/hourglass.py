"""Shim: put this directory ahead of the reference checkout on PYTHONPATH and the reference's own `train.py` /
`evaluate.py` / `export.py` (`from hourglass import StackedHourglass`, train.py:14) get the B200 network."""
from real_time_helmet_detection_b200.hourglass import *          # noqa: F401,F403
from real_time_helmet_detection_b200.hourglass import (StackedHourglass, Hourglass, Residual, Convolution,  # noqa: F401
                                                       PreLayer, Neck, Head, Pool, Activation)
