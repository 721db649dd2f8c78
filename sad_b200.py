"""Import alias: the package directory is named `3dsad-main_b200/` (not a valid Python
identifier), so `import sad_b200` loads it from there under this name."""
import importlib.util
import os
import sys

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3dsad-main_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_path, "__init__.py"), submodule_search_locations=[_path])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
