"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference (RudyMorel/shadowing).

Imports the unmodified reference modules from /root/reference with (i) a name-only stub of
the un-vendored `scatspectra` dependency (`path_shadowing.py:9`) and (ii) a synthetic parent
package so `shadowing/__init__.py:1-4` (matplotlib, PDV) is skipped.  Sources: /root/reference (build container) or
`baseline/_ref` (the same package pip-installed by the builder; it travels to the GPU box).  Used by
`tests/gen_golden.py` to produce the committed fixtures under `tests/golden/`, by the pin tests, and by
bench.py's CPU legs to time the reference's own `cuda=False` path beside the GPU.
Nothing in the product package imports this file.
"""
import importlib
import sys
import types
from pathlib import Path

# the build container mounts the reference at /root/reference; `baseline/_ref` holds the same unmodified
# package installed with pip (git-ignored, but it travels to the GPU box with the snapshot)
_CANDIDATES = [Path("/root/reference"), Path(__file__).resolve().parents[1] / "baseline" / "_ref"]
REF_ROOT = next((c for c in _CANDIDATES if (c / "shadowing" / "path_shadowing" / "path_shadowing.py").is_file()),
                _CANDIDATES[0])


def available() -> bool:
    return (REF_ROOT / "shadowing" / "path_shadowing" / "path_shadowing.py").is_file()


def load(softmax_cls=None, uniform_cls=None, proba_base=None):
    """Return the reference's `path_shadowing`, `path_embedding`, `path_distance`,
    `statistics` modules (unmodified source, torch-CPU).  `Softmax`/`Uniform` are not in the
    reference tree (scatspectra v2.0.2, README.md:21); pass stand-ins to exercise
    `predict_from_paths`, otherwise name-only placeholders are installed."""
    if not available():
        raise RuntimeError("live reference not present (expected in the build container only)")
    if "scatspectra" not in sys.modules or getattr(sys.modules["scatspectra"], "_psh_stub", False):
        stub = types.ModuleType("scatspectra")
        stub._psh_stub = True
        for name in ("TimeSeriesDataset", "PriceData", "windows"):
            setattr(stub, name, type(name, (), {}))
        stub.DiscreteProba = proba_base or type("DiscreteProba", (), {})
        stub.Softmax = softmax_cls or type("Softmax", (), {})
        stub.Uniform = uniform_cls or type("Uniform", (), {})
        sys.modules["scatspectra"] = stub
    if "shadowing" not in sys.modules or not hasattr(sys.modules["shadowing"], "_psh_synthetic"):
        pkg = types.ModuleType("shadowing")
        pkg.__path__ = [str(REF_ROOT / "shadowing")]
        pkg._psh_synthetic = True
        sys.modules["shadowing"] = pkg
        # drop cached submodules from an earlier load with other stand-ins
        for m in [m for m in sys.modules if m.startswith("shadowing.")]:
            del sys.modules[m]
    ps = importlib.import_module("shadowing.path_shadowing.path_shadowing")
    pe = importlib.import_module("shadowing.path_shadowing.path_embedding")
    pd = importlib.import_module("shadowing.path_shadowing.path_distance")
    st = importlib.import_module("shadowing.statistics")
    return types.SimpleNamespace(path_shadowing=ps, path_embedding=pe, path_distance=pd, statistics=st)
