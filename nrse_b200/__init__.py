"""Import alias: ``nrse_b200`` -> the package directory ``noise-robust-speech-embedding_b200/``.

The package directory carries the repository's hyphenated name, which Python cannot
import directly.  This shim loads it under the importable name ``nrse_b200`` so that
``import nrse_b200.models.byol`` etc. resolve to files in that directory.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
    "noise-robust-speech-embedding_b200",
)
_spec = importlib.util.spec_from_file_location(
    "nrse_b200",
    os.path.join(_PKG_DIR, "__init__.py"),
    submodule_search_locations=[_PKG_DIR],
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nrse_b200"] = _mod
_spec.loader.exec_module(_mod)
