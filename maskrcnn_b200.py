"""Import shim: loads the package directory ``mask-rcnn-coreml_b200/`` as module ``maskrcnn_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mask-rcnn-coreml_b200")
_spec = importlib.util.spec_from_file_location("maskrcnn_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["maskrcnn_b200"] = _mod
_spec.loader.exec_module(_mod)
