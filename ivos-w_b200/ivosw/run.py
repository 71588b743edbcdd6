"""Launcher: run an unmodified IVOS-W entry script on the B200 path.

    cd /path/to/IVOS-W
    PYTHONPATH=/path/to/repo/ivos-w_b200 python -m ivosw.run eval_agent_manet.py with setting=wild dataset=davis method=ours

Equivalent to ``python eval_agent_manet.py with ...`` (README.md:64 of the reference) except that
``ivosw.hook.install()`` runs first, so that ``models.agent``, ``models.assessment`` and ``utils.utils_agent`` are the
drop-in modules and ``utils.utils_manet`` / ``utils.utils_atnet`` get their round wrappers replaced (ivosw/hook.py).
The script runs as ``__main__`` with its own directory at ``sys.path[0]`` and ``sys.argv[0]`` set to its path, exactly
as the interpreter would start it.  ``--ivosw-report`` (before the script path) prints which modules were substituted
when the script ends.
"""
import os
import runpy
import sys

from . import hook


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    want_report = False
    while argv and argv[0].startswith("--ivosw-"):
        flag = argv.pop(0)
        if flag == "--ivosw-report":
            want_report = True
        else:
            raise SystemExit("ivosw.run: unknown option %s" % flag)
    if not argv:
        raise SystemExit("usage: python -m ivosw.run [--ivosw-report] /path/to/IVOS-W/eval_agent_{manet,atnet,ipn}.py [args ...]")
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit("ivosw.run: no such script: %s" % script)
    # what `python script.py` does: the script's directory first on sys.path, argv[0] = the script
    sys.path.insert(0, os.path.dirname(script))
    sys.argv = [argv[0]] + argv[1:]
    hook.install()
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        if want_report:
            import json
            print("ivosw.run: module substitution report\n" + json.dumps(hook.report(), indent=1, default=str), file=sys.stderr)


if __name__ == "__main__":
    main()
