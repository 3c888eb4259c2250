"""Imports the product package.  Its directory is named after the reference ("broadphase-rs_b200"),
which is not a valid Python identifier, so it is registered as module `broadphase_rs_b200`."""
import importlib.util
import os
import sys

NAME = "broadphase_rs_b200"
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "broadphase-rs_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(
        NAME, os.path.join(ROOT, "__init__.py"), submodule_search_locations=[ROOT])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
