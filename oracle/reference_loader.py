"""Import the UNMODIFIED reference (LuletterSoul/RP-Style-Transfer) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) pin `oracle/restate.py` against the real
reference functions and (b) generate the golden vectors under `tests/golden/`.  It only works in
the development container: `/root/reference` does not exist on the GPU box, so nothing in
`-m gpu` tests, `smoke()` or `bench.py` may call `load_reference()`.

The reference does not import on a modern stack: `network/base.py:2` imports
`numpy.lib.arraypad` (removed in numpy 2), `network/adain_rp.py:6,8` / `network/sanet.py:7-8`
import seaborn / matplotlib, and `utils/mst.py:3` imports PyMaxflow.  None of those is on the
hot path, so empty stub modules are injected for exactly those names.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RPST_REFERENCE_ROOT", "/root/reference")

_STUBS = {
    "numpy.lib.arraypad": {"pad": None},
    "seaborn": {},
    "matplotlib": {},
    "matplotlib.pyplot": {},
    "maxflow": {},
    "maxflow.fastmin": {"aexpansion_grid": None},
}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "network"))


def load_reference():
    """Return the reference's `network` package (imported once, under the stubs above)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (container-only helper)")
    if "network" in sys.modules and getattr(sys.modules["network"], "__rpst_reference__", False):
        return sys.modules["network"]
    for name, attrs in _STUBS.items():
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            if parent in sys.modules:
                setattr(sys.modules[parent], child, mod)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    net = importlib.import_module("network")
    net.__rpst_reference__ = True
    return net
