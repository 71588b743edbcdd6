"""ivosw — host side of the B200-native IVOS-W frame-scoring path.

    from ivosw import Engine, get_engine

The CUDA library (lib/libivosw_b200.so, C ABI in include/ivosw_b200.h) is loaded
eagerly by ``ivosw._lib``; importing ``ivosw.engine`` without it raises.
``ivosw.arch`` / ``ivosw.synth`` are pure-Python helpers (layer tables, seeded
synthetic inputs) and import without the library.
"""
from . import arch  # noqa: F401

__all__ = ["arch", "Engine", "get_engine"]


def __getattr__(name):
    if name in ("Engine", "get_engine", "pack_brain", "pack_assess"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
