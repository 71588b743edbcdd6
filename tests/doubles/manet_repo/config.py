"""TEST DOUBLE for MANet's config module (the reference reaches utils/config_manet/config.py through sys.path and reads
cfg.KNNS; the real file raises without CUDA, SURVEY A.Q11)."""
from types import SimpleNamespace

cfg = SimpleNamespace(KNNS=1, IS_TEST_DOUBLE=True)
