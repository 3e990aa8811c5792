"""Import shim: the product package lives in the directory ``mswe-gnn_b200/`` (a name Python
cannot import directly because of the hyphen).  ``import mswe_gnn_b200`` loads that directory
as a regular package (sub-modules resolve inside it), so user code can write
``from mswe_gnn_b200.models.gnn import MSGNN`` exactly like ``from models.gnn import MSGNN`` in
the reference."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mswe-gnn_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
