"""LRP rule and composite descriptors -- the configuration surface of the LRP pass.

The reference configures its LRP pass with zennit 0.5.1 objects: lists of ``(layer_names, rule)``
wrapped in ``NameMapComposite(name_map=..., canonizers=[SequentialMergeBatchNorm()])``
(utils/constants.py:27-51, drsa/cluster/getdrsadata.py:81-114).  zennit executes the rules through
Python hooks and ``torch.autograd``; here the same names are plain descriptors that the CUDA engine
(``lrp_engine``) reads -- the arithmetic of each rule lives in ``libdrsa_b200.so``.
Constructor arguments and defaults follow zennit.rules / zennit.composites / zennit.canonizers.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

__all__ = ["Rule", "Epsilon", "Gamma", "ZPlus", "WSquare", "Flat", "Pass", "NameMapComposite", "Composite",
           "SequentialMergeBatchNorm"]


class Rule:
    """Base descriptor.  ``kind`` selects the kernel path, ``stabilizer`` the epsilon of stabilize()."""
    kind = "none"

    def __init__(self, stabilizer: float = 1e-6):
        self.stabilizer = float(stabilizer)

    def copy(self):
        import copy
        return copy.copy(self)

    def __repr__(self):
        args = ", ".join(f"{k}={v}" for k, v in vars(self).items())
        return f"{type(self).__name__}({args})"


class Epsilon(Rule):
    """R_in = x * W^T (R_out / stabilize(Wx + b, epsilon))."""
    kind = "epsilon"

    def __init__(self, epsilon: float = 1e-6):
        super().__init__(stabilizer=epsilon)
        self.epsilon = float(epsilon)


class Gamma(Rule):
    """Positive weights emphasised by gamma.  Applied to layers with non-negative input, where
    zennit's generalised 4+1-pass rule collapses to one modified forward and one backward-data pass
    with W' = W + gamma*max(W,0), b' = b + gamma*max(b,0)."""
    kind = "gamma"

    def __init__(self, gamma: float = 0.25, stabilizer: float = 1e-6):
        super().__init__(stabilizer=stabilizer)
        self.gamma = float(gamma)


class ZPlus(Rule):
    """W' = max(W,0), b' = max(b,0) (the gamma -> infinity limit) for non-negative inputs."""
    kind = "zplus"


class WSquare(Rule):
    """Input replaced by ones, parameters squared, no input factor."""
    kind = "wsquare"


class Flat(Rule):
    """Input replaced by ones, parameters replaced by ones, no input factor."""
    kind = "flat"


class Pass(Rule):
    """Relevance passes through unchanged."""
    kind = "pass"

    def __init__(self):
        super().__init__(stabilizer=0.0)


class SequentialMergeBatchNorm:
    """Canonizer descriptor: fold every BatchNorm that directly follows a Conv2d / Linear into that
    layer's weight and bias for the duration of the pass (model must be in eval mode)."""


class Composite:
    def __init__(self, canonizers: Iterable = ()):  # noqa: D401
        self.canonizers = list(canonizers or [])

    def rule_for(self, name: str):
        return None

    @property
    def merges_batchnorm(self) -> bool:
        return any(isinstance(c, SequentialMergeBatchNorm) for c in self.canonizers)


class NameMapComposite(Composite):
    """Rules assigned by module name: ``name_map = [(['features.0'], WSquare(...)), ...]``."""

    def __init__(self, name_map: Sequence[Tuple[List[str], Rule]], canonizers: Iterable = ()):
        super().__init__(canonizers)
        self.name_map = list(name_map)
        self._by_name = {}
        for names, rule in self.name_map:
            for n in names:
                self._by_name[n] = rule

    def rule_for(self, name: str):
        return self._by_name.get(name)
