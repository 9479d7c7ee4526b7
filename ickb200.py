"""
Import shim: the product package lives in ``image-captioning-with-external-knowledge_b200/`` (a directory name that
is not a Python identifier).  ``import ickb200`` loads that directory as the package ``ickb200``.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "image-captioning-with-external-knowledge_b200")
_spec = importlib.util.spec_from_file_location(
    "ickb200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ickb200"] = _mod
_spec.loader.exec_module(_mod)
