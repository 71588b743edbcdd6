"""TEST DOUBLE for ATNet's config.py (eval_agent_atnet.py:28, 88)."""


class Config(object):
    test_propagation_proportion = 1.0
    scribble_dilation_param = 3
    mean, var = 0.45, 0.22
    davis_dataset_dir = ''
    test_propth = 0.5
