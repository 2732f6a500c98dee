"""Import shim: the product package lives in ``relation-networks-for-object-detection_b200/`` (a directory name Python
cannot import directly); ``import relnet_b200`` loads it under that name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'relation-networks-for-object-detection_b200')
_spec = importlib.util.spec_from_file_location('relnet_b200', os.path.join(_dir, '__init__.py'),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['relnet_b200'] = _mod
_spec.loader.exec_module(_mod)
