"""TEST DOUBLE for MANet's networks/deeplab.py (external repo lightas/CVPR2020_MANet, absent)."""


class DeepLab(object):
    def __init__(self, *a, **k):
        pass
