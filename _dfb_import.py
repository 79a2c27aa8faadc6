"""Registers the package directory `digital-filtering_b200/` (a name Python cannot import as
written) as the module `digital_filtering_b200`.  `import _dfb_import` once, then
`import digital_filtering_b200`."""
import importlib.util
import os
import sys

_NAME = "digital_filtering_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "digital-filtering_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
