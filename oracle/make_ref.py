"""Recipe for ``oracle/_ref/``: a byte copy of the reference's Python sources for the hot path, so that the UNMODIFIED
reference code can be timed on the GPU box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``).

TEST / BASELINE INFRASTRUCTURE ONLY.  /root/reference exists only in the build container; ``oracle/_ref/`` is git-ignored
(reference sources never enter this repository's history) but travels to the GPU box with the snapshot.  Run by
``__graft_entry__.build()`` whenever the reference tree is present; a no-op otherwise.

The reference is pure Python -- there is nothing to compile.  Files are copied verbatim (sha256 recorded in
``oracle/_ref/MANIFEST.json``); the typo at drsa.py:4 (``from pathilib import Path``) is NOT patched: loaders alias the
module name ``pathilib`` to ``pathlib`` instead."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("DRSA_REFERENCE_ROOT", "/root/reference")


# the files of the hot path and what they import (nothing else of the reference is copied)
FILES = ["cxai/__init__.py", "cxai/model/__init__.py", "cxai/model/create_model.py", "cxai/xai/__init__.py",
         "cxai/xai/drsa/__init__.py", "cxai/xai/drsa/drsa.py", "cxai/xai/drsa/preprocessing.py",
         "cxai/xai/explain/__init__.py", "cxai/xai/explain/attribute.py", "cxai/utils/__init__.py",
         "cxai/utils/constants.py", "cxai/utils/dataloading.py", "cxai/utils/sound.py", "cxai/utils/utilities.py"]


def make() -> bool:
    src = os.path.join(REFERENCE_ROOT, "cxai")
    if not os.path.isdir(src):
        return os.path.isfile(os.path.join(REF_OUT, "MANIFEST.json"))
    shutil.rmtree(os.path.join(REF_OUT, "cxai"), ignore_errors=True)
    manifest = {}
    for rel in FILES:
        if not os.path.isfile(os.path.join(REFERENCE_ROOT, rel)):
            continue
        dst = os.path.join(REF_OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE_ROOT, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(REF_OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "sharckhai/drsa-audio (verbatim copies, see oracle/make_ref.py)", "files": manifest}, fh, indent=1)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(REF_OUT, "cxai", "xai", "drsa", "drsa.py"))


def load_drsa():
    """The reference's ``cxai/xai/drsa/drsa.py`` from ``oracle/_ref`` executed as-is under a private module name."""
    import importlib.util
    import pathlib
    import sys
    sys.modules.setdefault("pathilib", pathlib)          # drsa.py:4 typo
    spec = importlib.util.spec_from_file_location("_reference_drsa_copy", os.path.join(REF_OUT, "cxai/xai/drsa/drsa.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("oracle/_ref ready:", make())
