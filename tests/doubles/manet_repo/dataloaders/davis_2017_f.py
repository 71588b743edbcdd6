"""TEST DOUBLE for MANet's dataloaders/davis_2017_f.py"""


class DAVIS2017_Feature_Extract(object):
    def __init__(self, *a, **k):
        pass
