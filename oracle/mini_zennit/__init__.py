"""mini-zennit: a restatement of the part of ``zennit==0.5.1`` that the reference calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Why it exists: every line of LRP arithmetic of the reference lives in the third-party package zennit
(reference ``requirements.txt:23`` pins 0.5.1), which is neither vendored under /root/reference nor installed nor
obtainable in this environment (no network).  ``install()`` registers this package as ``sys.modules['zennit']`` (and its
sub-modules ``core, rules, composites, canonizers, attribution, types``) so that the reference's OWN stage-1 code --
``cxai/xai/drsa/preprocessing.py:92-176``, ``cxai/xai/explain/attribute.py:12-160``,
``cxai/xai/explain/explainer.py:15-203``, ``cxai/model/modify_model.py:4-123``, ``cxai/utils/constants.py:27-51`` --
imports and runs UNMODIFIED on top of it (``oracle/gen_golden_lrp.py``).

PARITY STATUS: the orchestration above it is the reference's; the rule arithmetic below is a RESTATEMENT of zennit
0.5.1's published semantics (hook plumbing through ``grad_fn.register_hook`` on an identity node, ``BasicHook.backward``
with input / parameter / output modifiers, gradient mapper and reducer; rules Epsilon, Gamma, ZPlus, AlphaBeta, ZBox,
WSquare, Flat, Pass, Norm; ``SequentialMergeBatchNorm``; ``Composite`` family; ``Gradient`` attributor), written
from knowledge of that release -- "parity unpinned" for exactly this layer until the real package (or vectors from
it) can be brought in.  One detail could not be settled from memory and is a module-level switch:
``rules.FLAT_ZERO_BIAS`` (whether ``Flat`` sets the bias to one or drops it).
"""
import sys

from . import core, rules, composites, canonizers, attribution, types  # noqa: F401

__version__ = "0.5.1+restated"


def install() -> None:
    """Register this package under the name ``zennit`` (only if the real package is absent)."""
    try:
        import importlib.util
        if importlib.util.find_spec("zennit") is not None and "zennit" not in sys.modules:
            raise RuntimeError("a real zennit is installed: use it instead of the restatement")
    except (ImportError, ValueError):
        pass
    me = sys.modules[__name__]
    sys.modules["zennit"] = me
    for sub in ("core", "rules", "composites", "canonizers", "attribution", "types"):
        sys.modules[f"zennit.{sub}"] = getattr(me, sub)
