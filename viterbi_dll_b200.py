"""Import shim: the package directory is named ``viterbi.dll_b200`` (with a dot), which the
normal import statement cannot spell.  ``import viterbi_dll_b200 as vb`` loads it by path."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "viterbi.dll_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
