"""Import shim: the package directory is named `faster-rcnn.torch_b200` (not a valid Python identifier), so it is
loaded here under the module name `frcnn_b200`.  `import frcnn_b200` from the repo root gives the package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "faster-rcnn.torch_b200")
_spec = importlib.util.spec_from_file_location("frcnn_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["frcnn_b200"] = _mod
_spec.loader.exec_module(_mod)
