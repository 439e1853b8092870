"""Run one of the reference's scripts, unmodified, on the GPU drop-in modules.

    python -m ia2c_b200.launcher /path/to/IA2C/a2c_org_test.py [--set n_updates=100]
    python -m ia2c_b200.launcher /path/to/IA2C/ia2c.py --set NUM_EPISODES=200 --set n_envs=4096

The reference's scripts (thinclab/IA2C: ia2c.py:20-23,33-38; a2c_org_test.py:16-19) obtain everything through four
module names — ``Org``, ``ac_nets``, ``belief_filter`` and ``gymnasium`` (entry point ``"Org:Org"``).  This launcher
puts ``ia2c_b200/compat`` (and ``ia2c_b200/compat_gym`` when the real gymnasium is not installed) in FRONT of the
module search path, so those names resolve to the drop-in modules instead of the files that sit next to the
script, and then executes the script's source as ``__main__``.  ``--set NAME=VALUE`` rewrites a module-level
constant assignment (``NAME = 123`` at the start of a line) in the in-memory source — the scripts have no command
line of their own (ia2c.py:25-31,40; a2c_org_test.py:22-25) — nothing is written to disk.
"""
from __future__ import annotations

import argparse
import os
import re
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))


def dropin_paths():
    """Directories to put in front of sys.path, most specific last (so that it ends up first)."""
    paths = []
    try:
        import gymnasium  # noqa: F401  the real package: compat/Org.py routes make_vec("Org-v0") to the GPU vector env
    except Exception:
        paths.append(os.path.join(_PKG, "compat_gym"))
    paths.append(os.path.join(_PKG, "compat"))
    return paths


def install_dropins():
    for p in dropin_paths():
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


def substitute(src, settings, fname="<script>"):
    for name, value in settings.items():
        src, n = re.subn(rf"^{re.escape(name)}\s*=\s*[^\n#]+?(?=\s*(?:#|$))", f"{name} = {value}", src, count=1, flags=re.M)
        if n != 1:
            raise ValueError(f"{fname}: no module-level assignment of {name!r} to rewrite")
    return src


def run_script(path, settings=None, namespace=None):
    """Execute ``path`` as __main__ on the drop-in modules; returns the script's global namespace."""
    install_dropins()
    path = os.path.abspath(path)
    src = substitute(open(path).read(), settings or {}, os.path.basename(path))
    ns = {"__name__": "__main__", "__file__": path}
    if namespace:
        ns.update(namespace)
    exec(compile(src, path, "exec"), ns)
    return ns


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("script")
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE", help="rewrite a module-level constant of the script")
    a = ap.parse_args(argv)
    settings = {}
    for item in a.set:
        if "=" not in item:
            ap.error(f"--set expects NAME=VALUE, got {item!r}")
        k, v = item.split("=", 1)
        settings[k.strip()] = v.strip()
    run_script(a.script, settings)


if __name__ == "__main__":
    main()
