"""TEST DOUBLE for the checkout's utils/utils_manet.py: load_network / preprocess must survive the patch;
get_results / rough_ROI must be REPLACED by ivosw.hook (the ones below raise)."""
from config import cfg  # noqa: F401  (MANet's config module; tests/doubles/third_party.py provides a stand-in)

IS_CHECKOUT_ORIGINAL = True


def load_network(net, pretrained_dict):
    model_dict = net.state_dict()
    model_dict.update({k: v for k, v in pretrained_dict.items() if k in model_dict})
    net.load_state_dict(model_dict)


def preprocess(db_root_dir, seqs, seq_list_file):
    return {s: [1] for s in seqs}


def rough_ROI(ref_scribble_labels):
    raise AssertionError("checkout's rough_ROI called: the hook did not patch utils.utils_manet")


def get_results(*a, **k):
    raise AssertionError("checkout's get_results called: the hook did not patch utils.utils_manet")
