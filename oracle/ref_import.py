"""Load the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so this
module is used solely by ``oracle/gen_golden.py`` (fixture generation) and by tests
that skip when the reference tree is absent.

`drsa.py` has a typo at line 4 (``from pathilib import Path``); we alias the missing
module name to ``pathlib`` and execute the file as-is under a private module name so
it does not collide with this repo's own ``cxai`` package.
"""
from __future__ import annotations

import importlib.util
import os
import pathlib
import sys

REFERENCE_ROOT = os.environ.get("DRSA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cxai/xai/drsa/drsa.py"))


def load_reference_drsa():
    """Returns the reference's ``cxai.xai.drsa.drsa`` module object (drsa.py:1-300)."""
    sys.modules.setdefault("pathilib", pathlib)
    path = os.path.join(REFERENCE_ROOT, "cxai/xai/drsa/drsa.py")
    spec = importlib.util.spec_from_file_location("_reference_drsa", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
