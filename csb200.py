"""Loader for the package directory ``cellscopes.jl_b200/`` (its name has a dot, so a plain
``import`` cannot reach it).  ``import csb200 as cs`` gives the package module itself."""
import importlib.util
import os
import sys

_NAME = "cellscopes_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cellscopes.jl_b200")

if _NAME in sys.modules:
    _pkg = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _pkg = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _pkg
    _spec.loader.exec_module(_pkg)

sys.modules[__name__] = _pkg
