"""Loader for the product package.  Its directory is named `fluid-rs_b200` (not an importable
identifier), so entry points call `fluidpkg.load()` and get it as module `fluid_rs_b200`."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "fluid-rs_b200"


def load():
    if "fluid_rs_b200" in sys.modules:
        return sys.modules["fluid_rs_b200"]
    spec = importlib.util.spec_from_file_location(
        "fluid_rs_b200", PKG_DIR / "__init__.py", submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fluid_rs_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
