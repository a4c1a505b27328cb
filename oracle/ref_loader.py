"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference solver.

Executes /root/reference/RPModule/rpmodule.py in memory, exactly as SURVEY.md
section 8(c) describes:

  * the file does not parse as shipped (RPModule/rpmodule.py:342-343 read
    ``dataS['feat'] / FEAT_SCALING.``); the stale .pyc of the previous revision
    and ``FEAT_SCALING = 100`` at rpmodule.py:327 pin the intended semantics to
    ``feat / 100``; the two trailing dots are removed in memory,
  * ``matplotlib`` / ``open3d`` (unused by the solver; rpmodule.py:10,
    rputil.py:4) are replaced by empty stub modules when absent.

This works where /root/reference exists (the build container) or where baseline/_ref was installed from it
(oracle/install_reference.py; travels to the GPU box, git-ignored).  It is used
by tests/golden/make_golden.py to produce the committed golden vectors and by
the ``not gpu`` tests to re-pin the numpy restatement when the tree is present.
and by bench.py's CPU arm (cpu_baseline.kind == "reference").  Nothing in the product path may import this module.
"""
import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """$RP_REFERENCE_ROOT, else the build container's read-only tree, else the offline install that travels to the GPU
    box (baseline/_ref, made by oracle/install_reference.py; unmodified files)."""
    cands = [os.environ.get("RP_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "RPModule", "rpmodule.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "RPModule", "rpmodule.py"))


def _stub(name):
    if name in sys.modules:
        return
    try:
        if importlib.util.find_spec(name) is not None:
            return
    except (ImportError, ValueError):
        pass
    mod = types.ModuleType(name)
    mod.__path__ = []
    sys.modules[name] = mod


_cached = None


def load_reference_rpmodule():
    """Return the reference ``RPModule.rpmodule`` module object (patched in memory)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "open3d"):
        _stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # The repo root carries drop-in packages named RPModule/ and model/ too; make
    # sure the names below resolve to the *reference* while we exec it.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "RPModule" or k.startswith("RPModule.") or k == "util" or k == "utils"
             or k.startswith("utils.")}
    try:
        pkg = types.ModuleType("RPModule")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "RPModule")]
        sys.modules["RPModule"] = pkg
        path = os.path.join(REFERENCE_ROOT, "RPModule", "rpmodule.py")
        with open(path) as fh:
            src = fh.read()
        bad = "/ FEAT_SCALING.\n"
        assert src.count(bad) == 2, "reference changed: expected two '/ FEAT_SCALING.' lines"
        src = src.replace(bad, "/ FEAT_SCALING\n")
        mod = types.ModuleType("RPModule.rpmodule")
        mod.__file__ = path
        mod.__package__ = "RPModule"
        sys.modules["RPModule.rpmodule"] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        ref_mod = mod
        ref_util = sys.modules.get("RPModule.rputil")
    finally:
        # leave the reference modules registered under private names only
        for k in [k for k in sys.modules if k == "RPModule" or k.startswith("RPModule.")
                  or k == "util" or k == "utils" or k.startswith("utils.")]:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    ref_mod._rputil = ref_util
    _cached = ref_mod
    return ref_mod


def reference_opts(*a, **k):
    """Instance of the reference's ``rputil.opts`` (RPModule/rputil.py:11-22)."""
    return load_reference_rpmodule()._rputil.opts(*a, **k)
