"""pnec_b200 — B200-native PNEC frame-pair rotation/translation refinement.

Drop-in for the Ceres-backed refinement of tum-vision/pnec
(pnec::rel_pose_estimation::PNEC::CeresSolver & co, pypnec.pyceres / pyceresnec):
hand-written sm_100a CUDA behind a C-ABI (include/pnec_b200.h).

    pnec_b200.api          ctypes binding of the C-ABI (batched solve / eval / cost)
    pnec_b200.pypnec       compiled pybind11 module with the reference's Python API
    pnec_b200.synthetic    the reference's simulation data distribution, vectorised
    pnec_b200.distributed  contiguous batch sharding + pose gather (one process per GPU)

There is no CPU fallback anywhere in this package.
"""
import importlib
import os
import sys

__version__ = "0.1"

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def __getattr__(name):
    if name == "pypnec":
        if _LIB_DIR not in sys.path:
            sys.path.insert(0, _LIB_DIR)
        try:
            mod = importlib.import_module("pypnec")
        except ImportError as e:  # fail loudly: the extension is the product
            raise ImportError(
                "pnec_b200.pypnec is not built: run `python -c 'import __graft_entry__ as g; g.build()'`"
            ) from e
        globals()["pypnec"] = mod
        return mod
    if name in ("api", "synthetic", "distributed"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
