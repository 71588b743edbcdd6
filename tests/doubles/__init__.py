"""Test doubles (see README.md) and the helper that activates the drop-in hook over the double checkout."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CHECKOUT = os.path.join(HERE, "checkout")


def purge(prefixes=("models", "utils", "datasets", "libs", "ivosw_dropin", "config", "networks", "dataloaders")):
    for m in [k for k in sys.modules if any(k == p or k.startswith(p + ".") for p in prefixes)]:
        del sys.modules[m]


def activate(checkout=CHECKOUT):
    """sys.path[0] = checkout (what `python script.py` does), hook installed.  Returns the hook module."""
    from ivosw import hook
    hook.uninstall()
    purge()
    if checkout in sys.path:
        sys.path.remove(checkout)
    sys.path.insert(0, checkout)
    hook.install()
    return hook


def deactivate(checkout=CHECKOUT):
    from ivosw import hook
    hook.uninstall()
    _only_repo(None)
    purge()
    if checkout in sys.path:
        sys.path.remove(checkout)


def _only_repo(which):
    """exactly one of the external-repo doubles on sys.path (each has its own top-level `config` / `networks`, as the
    real MANet / ATNet checkouts do: an entry script only ever appends one of them)"""
    for extra in ("manet_repo", "atnet_repo", "ipn_repo"):
        p = os.path.join(HERE, extra)
        while p in sys.path:
            sys.path.remove(p)
    purge(("config", "networks", "dataloaders", "libs"))
    if which:
        sys.path.insert(1, os.path.join(HERE, which))


def load_dropin(with_manet=False, with_atnet=False):
    """The drop-in modules as an entry script sees them: imported under the reference's names, through the hook, over
    the double checkout.  Returns a namespace (A = models.agent, S = models.assessment, U = utils.utils_agent,
    M = utils.utils_manet, AT = utils.utils_atnet).  With with_atnet the ATNet repo double stays on sys.path afterwards
    (tests import its stand-in network from there)."""
    from types import SimpleNamespace
    activate()
    _only_repo(None)
    import models.agent as A
    import models.assessment as S
    import utils.misc as misc
    import utils.utils_agent as U
    ns = SimpleNamespace(A=A, S=S, U=U, misc=misc, M=None, AT=None)
    if with_manet:
        _only_repo("manet_repo")
        import utils.utils_manet as M
        ns.M = M
    if with_atnet:
        _only_repo("atnet_repo")
        import utils.utils_atnet as AT
        ns.AT = AT
    return ns
