"""TEST DOUBLE for IPN's model.py (external repo zyy-cn/IPN, absent): import surface only (eval_agent_ipn.py:27)."""


class model(object):
    def __init__(self, *a, **k):
        pass
