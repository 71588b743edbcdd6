"""Module substitution for an UNMODIFIED IVOS-W checkout (the boundary of the reference is a
handful of Python imports in its entry scripts).

``install()`` puts one finder at the head of ``sys.meta_path``:

  substituted   ``models.agent``, ``models.assessment``, ``utils.utils_agent`` are loaded from
                ``ivos-w_b200/dropin/`` under the reference's module names (``eval_agent_manet.py:27-30``,
                ``eval_agent_atnet.py:22-26``, ``eval_agent_ipn.py``).  The parent packages ``models`` and ``utils``
                stay the checkout's own, so ``utils.misc``, ``models.momory_pool`` (the full ReplayMemory with its
                CSV mirror), ``utils.utils_ipn``, ``datasets`` ... resolve exactly as before.
  patched       ``utils.utils_manet`` and ``utils.utils_atnet`` import normally from the checkout (their
                ``load_network`` / ``preprocess`` and their own imports of ``config``, ``libs``, ``datasets`` are
                untouched); afterwards ``get_results`` + ``rough_ROI`` resp. ``run_VOS_singleiact`` are replaced by
                the drop-in versions, which keep calling the external VOS networks the way the reference does.

Why a finder and not ``PYTHONPATH``: Python puts the script's directory at ``sys.path[0]``, ahead of
``PYTHONPATH``, so a path entry can never shadow the checkout's own ``models/`` and ``utils/`` packages — and a
shadowing package would hide ``utils.misc`` ("What's weak" #2).

Use ``python -m ivosw.run /path/to/IVOS-W/eval_agent_manet.py with ...`` (ivosw/run.py), or call
``ivosw.hook.install()`` before the first ``import models`` / ``import utils``.
"""
import importlib.abc
import importlib.util
import os
import sys

DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")

SUBSTITUTED = {
    "models.agent": os.path.join("models", "agent.py"),
    "models.assessment": os.path.join("models", "assessment.py"),
    "utils.utils_agent": os.path.join("utils", "utils_agent.py"),
}
# reference module -> (drop-in file, attributes copied into the reference module after it has been imported)
PATCHED = {
    "utils.utils_manet": (os.path.join("utils", "utils_manet.py"), ("get_results", "rough_ROI")),
    "utils.utils_atnet": (os.path.join("utils", "utils_atnet.py"), ("run_VOS_singleiact",)),
}


def _load_private(fullname, relpath):
    """The drop-in file as a module of its own (``ivosw_dropin.<name>``), never registered under the reference's name."""
    name = "ivosw_dropin." + fullname.replace(".", "_")
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(DROPIN, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _PatchingLoader(importlib.abc.Loader):
    def __init__(self, inner, fullname):
        self.inner, self.fullname = inner, fullname

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)                      # the checkout's own module body, unmodified
        relpath, names = PATCHED[self.fullname]
        ours = _load_private(self.fullname, relpath)
        ours._ref = module                                  # external names (libs, datasets, DataLoader) come from here
        if hasattr(module, "cfg"):                          # MANet's config object (utils_manet.py:8)
            ours.cfg = module.cfg
        for n in names:
            setattr(module, "_reference_" + n, getattr(module, n, None))
            setattr(module, n, getattr(ours, n))
        module.__ivosw_patched__ = names


class DropinFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname in SUBSTITUTED:
            return importlib.util.spec_from_file_location(fullname, os.path.join(DROPIN, SUBSTITUTED[fullname]))
        if fullname in PATCHED:
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                spec = finder.find_spec(fullname, path, target)
                if spec is not None and spec.loader is not None:
                    spec.loader = _PatchingLoader(spec.loader, fullname)
                    return spec
        return None


_FINDER = None


def install():
    """Idempotent.  Must run before the entry script's first ``from models... / from utils...`` import."""
    global _FINDER
    if _FINDER is None:
        already = [m for m in list(SUBSTITUTED) + list(PATCHED) if m in sys.modules]
        if already:
            raise RuntimeError("ivosw.hook.install() came too late: %s already imported from the checkout" % already)
        _FINDER = DropinFinder()
        sys.meta_path.insert(0, _FINDER)
    return _FINDER


def uninstall():
    global _FINDER
    if _FINDER is not None and _FINDER in sys.meta_path:
        sys.meta_path.remove(_FINDER)
    _FINDER = None


def report():
    """Which modules of the running process come from the drop-in (for logs / tests)."""
    out = {}
    for name in SUBSTITUTED:
        m = sys.modules.get(name)
        out[name] = getattr(m, "__file__", None) if m else None
    for name in PATCHED:
        m = sys.modules.get(name)
        out[name] = {"file": getattr(m, "__file__", None), "patched": getattr(m, "__ivosw_patched__", ())} if m else None
    # modules of the checkout that must NOT be touched (a shadowing package used to hide utils.misc)
    out["kept"] = {name: getattr(sys.modules[name], "__file__", None)
                   for name in ("utils.misc", "models.momory_pool", "utils.utils_ipn", "datasets") if name in sys.modules}
    return out
