"""ORACLE — TEST INFRASTRUCTURE ONLY.

Loads the REAL reference voxeliser (/root/reference/utils.py) in the build container so that golden
vectors can be generated from the reference itself (tests/golden/make_voxel_golden.py).  utils.py imports
matplotlib, skimage, pymatgen and func_timeout at module top (utils.py:17-31) although density_matrix and
coordinate_grid only need numpy + scipy; those unrelated modules are stubbed in sys.modules.
/root/reference does not exist on the GPU box: nothing at test/bench run time may call this.
"""
from __future__ import annotations

import importlib.util
import sys
import types

REFERENCE_UTILS = "/root/reference/utils.py"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def load_reference_utils(path=REFERENCE_UTILS):
    def func_set_timeout(_seconds):
        return lambda f: f

    class FunctionTimedOut(Exception):
        pass

    _stub("func_timeout", FunctionTimedOut=FunctionTimedOut, func_set_timeout=func_set_timeout,
          func_timeout=lambda *a, **k: None)
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.feature", "pymatgen", "pymatgen.io",
                 "pymatgen.io.cif", "pymatgen.transformations", "pymatgen.transformations.standard_transformations"):
        try:
            __import__(name)
        except Exception:  # noqa: BLE001
            _stub(name)
    sys.modules["skimage.feature"].__dict__.setdefault("peak_local_max", None)
    sys.modules["pymatgen.io.cif"].__dict__.setdefault("CifParser", None)
    t = sys.modules["pymatgen.transformations.standard_transformations"].__dict__
    t.setdefault("AutoOxiStateDecorationTransformation", None)
    t.setdefault("OrderDisorderedStructureTransformation", None)
    spec = importlib.util.spec_from_file_location("icsg3d_reference_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
