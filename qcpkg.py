"""Registers the hyphen-named package directory `qchem-rs_b200/` as the module `qchem_rs_b200`."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
NAME = "qchem_rs_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    pkg_dir = ROOT / "qchem-rs_b200"
    spec = importlib.util.spec_from_file_location(NAME, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
