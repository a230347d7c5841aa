"""oracle/ref_loader.py — TEST INFRASTRUCTURE: import the UNMODIFIED reference ComA classes (snuvclab/coma utils/coma.py,
utils/coma_occupancy.py) from `oracle/_ref` (staged by oracle/make_ref.py) or, in the dev container, from /root/reference.

The reference's modules import each other as `utils.*`, the name of this repo's pickle-compat shim package. `load()`
therefore imports them with the reference directory first on sys.path and the repo's `utils*` modules temporarily out of
sys.modules, then restores both — the returned module objects keep their own references, so product and reference classes
can live in one process (tests compare them directly). open3d / trimesh / easydict are stubbed.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def reference_root():
    """oracle/_ref if staged; in the dev container (reference mounted) it is (re)staged on demand."""
    cand = os.path.join(_HERE, "_ref")
    if not os.path.exists(os.path.join(cand, "utils", "coma.py")):
        from . import make_ref
        make_ref.make()
    return cand if os.path.exists(os.path.join(cand, "utils", "coma.py")) else None


def available():
    return reference_root() is not None


def load():
    """-> namespace with ComA, ComA_Occupancy, get_aggregated_contact, coma (module), coma_occupancy (module), root."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules not staged: run `python oracle/make_ref.py` in the dev container")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")}
    stubs = {}
    for m in ("open3d", "trimesh", "easydict"):
        if m not in sys.modules:
            stubs[m] = sys.modules[m] = types.ModuleType(m)
    if "easydict" in stubs:
        stubs["easydict"].EasyDict = dict
    sys.path.insert(0, root)
    try:
        importlib.invalidate_caches()
        coma = importlib.import_module("utils.coma")
        occ = importlib.import_module("utils.coma_occupancy")
        assert os.path.abspath(coma.__file__).startswith(os.path.abspath(root)), coma.__file__
    finally:
        sys.path.remove(root)
        ref_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.")}
        sys.modules.update(saved)
    ns = types.SimpleNamespace(ComA=coma.ComA, ComA_Occupancy=occ.ComA_Occupancy, get_aggregated_contact=coma.get_aggregated_contact,
                               coma=coma, coma_occupancy=occ, root=root, _modules=ref_mods)
    _CACHE["ns"] = ns
    return ns
