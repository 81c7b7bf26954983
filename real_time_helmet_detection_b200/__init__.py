"""B200-native hot path of tyui592/Real_Time_Helmet_Detection.

Stacked-hourglass forward/backward, CenterNet decode + NMS and the focal / masked-L1 training losses as
hand-written sm_100a CUDA kernels behind the reference's own Python call signatures.
"""
__version__ = "0.1.0"
