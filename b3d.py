"""Import shim: registers the hyphenated package directory as the importable module `unet3d_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "segmentation-and-classification-of-brain-tumor-using-3d-unet_b200")

if "unet3d_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        "unet3d_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["unet3d_b200"] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules["unet3d_b200"]
