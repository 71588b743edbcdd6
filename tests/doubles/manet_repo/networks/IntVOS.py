"""TEST DOUBLE for MANet's networks/IntVOS.py — NOT MANet: returns synthetic logits with the return structure the
wrapper expects (utils/utils_manet.py:65-74, 92-107).  `logits` T x C x h x w is supplied by the test."""


class IntVOS(object):
    dynamic_seghead = None

    def __init__(self, cfg=None, feature_extracter=None, logits=None):
        self.logits = logits

    def int_seghead(self, **kw):
        return {kw["seq_names"][0]: self.logits[kw["frame_num"][0]][None]}, kw["local_map_dics"]

    def prop_seghead(self, *a, **kw):
        return ({kw["seq_names"][0]: self.logits[kw["frame_num"][0]][None]}, kw["global_map_tmp_dic"], kw["local_map_dics"])
