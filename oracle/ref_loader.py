"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *real* reference (garrekstemo/pistachio) from /root/reference/src so
that the oracle restatement (oracle/tmm_oracle.py) and the golden fixtures
(tests/golden/*.npz) can be pinned against it.  This only works in the build
container: /root/reference does not exist on the GPU box, so nothing under
tests -m gpu, smoke() or bench.py may call load_reference().

The reference imports `ruamel.yaml` (transfer_matrix.py:15), which is not
installed here; a PyYAML-backed stand-in module is registered in sys.modules
before the import.  Nothing in the reference is modified.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "pistachio", "transfer_matrix.py"))


def _install_ruamel_shim() -> None:
    try:
        import ruamel.yaml  # noqa: F401
        return
    except Exception:
        pass
    import yaml as _pyyaml

    class YAML:  # minimal surface used at transfer_matrix.py:22,375
        def __init__(self, *a, **k):
            pass

        def load(self, stream):
            return _pyyaml.safe_load(stream)

    pkg = types.ModuleType("ruamel")
    pkg.__path__ = []  # mark as package
    sub = types.ModuleType("ruamel.yaml")
    sub.YAML = YAML
    pkg.yaml = sub
    sys.modules.setdefault("ruamel", pkg)
    sys.modules.setdefault("ruamel.yaml", sub)


def load_reference():
    """Return the reference's `pistachio.transfer_matrix` module (unmodified)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    _install_ruamel_shim()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("pistachio.transfer_matrix")


def reference_default_yaml(name: str) -> str:
    return os.path.join(REFERENCE_SRC, "pistachio", "default", name)


def reference_csv(name: str) -> str:
    return os.path.join(REFERENCE_SRC, "pistachio", "data", "refractive_index_data", name)
