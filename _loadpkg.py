"""Import helper: the package directory is named `structuredgridcalc-boxframework_b200` (with a
hyphen, as the project layout prescribes), which `import` cannot spell.  load_package() imports it
under the module name `sgc_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "structuredgridcalc-boxframework_b200")


def load_package(name="sgc_b200"):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_build():
    spec = importlib.util.spec_from_file_location("sgc_b200_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
