"""zennit.canonizers (0.5.1), restated: batch-norm merging.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch

from .core import collect_leaves
from .types import BatchNorm, ConvolutionTranspose, Linear


class Canonizer:
    def apply(self, root_module):
        return []

    def register(self, *args, **kwargs):
        raise NotImplementedError

    def remove(self):
        raise NotImplementedError

    def copy(self):
        return self.__class__()


class MergeBatchNorm(Canonizer):
    """Folds a BatchNorm (eval mode) into the linear layer(s) in front of it and turns the BatchNorm into the identity;
    everything is restored by ``remove``."""
    linear_type = (Linear,)
    batch_norm_type = (BatchNorm,)

    def __init__(self):
        super().__init__()
        self.linears = None
        self.batch_norm = None
        self.linear_params = None
        self.batch_norm_params = None

    def register(self, linears, batch_norm):
        self.linears = linears
        self.batch_norm = batch_norm
        self.linear_params = [(linear.weight.data, None if linear.bias is None else linear.bias.data, linear.bias is None)
                              for linear in linears]
        self.batch_norm_params = {key: getattr(batch_norm, key).data
                                  for key in ('weight', 'bias', 'running_mean', 'running_var')}
        self.batch_norm_eps = batch_norm.eps
        self.merge_batch_norm(self.linears, self.batch_norm)

    def remove(self):
        for linear, (weight, bias, had_none) in zip(self.linears, self.linear_params):
            linear.weight.data = weight
            if had_none:
                linear.bias = None
            else:
                linear.bias.data = bias
        for key, value in self.batch_norm_params.items():
            getattr(self.batch_norm, key).data = value
        self.batch_norm.eps = self.batch_norm_eps

    @staticmethod
    def merge_batch_norm(modules, batch_norm):
        denominator = (batch_norm.running_var + batch_norm.eps) ** .5
        scale = (batch_norm.weight / denominator)
        for module in modules:
            original_weight = module.weight.data
            if module.bias is None:
                module.bias = torch.nn.Parameter(
                    torch.zeros(original_weight.shape[0], device=original_weight.device, dtype=original_weight.dtype))
            original_bias = module.bias.data
            if isinstance(module, ConvolutionTranspose):
                index = (None, slice(None), *((None,) * (original_weight.ndim - 2)))
            else:
                index = (slice(None), *((None,) * (original_weight.ndim - 1)))
            module.weight.data = (original_weight * scale[index])
            module.bias.data = (original_bias - batch_norm.running_mean) * scale + batch_norm.bias
        batch_norm.running_mean.data = torch.zeros_like(batch_norm.running_mean.data)
        batch_norm.running_var.data = torch.ones_like(batch_norm.running_var.data)
        batch_norm.bias.data = torch.zeros_like(batch_norm.bias.data)
        batch_norm.weight.data = torch.ones_like(batch_norm.weight.data)
        # zennit sets eps = 0.; torch >= 2.9 rejects a non-positive eps, and 1 + 1e-30 == 1 exactly in fp32 and fp64,
        # so the batch norm is the same bit-exact identity
        batch_norm.eps = 1e-30


class SequentialMergeBatchNorm(MergeBatchNorm):
    """Merges every BatchNorm that directly follows a linear layer in the order of the leaf modules."""

    def apply(self, root_module):
        instances = []
        last_leaf = None
        for leaf in collect_leaves(root_module):
            if isinstance(last_leaf, self.linear_type) and isinstance(leaf, self.batch_norm_type):
                instance = self.copy()
                instance.register((last_leaf,), leaf)
                instances.append(instance)
            last_leaf = leaf
        return instances


class NamedMergeBatchNorm(MergeBatchNorm):
    def __init__(self, name_map):
        super().__init__()
        self.name_map = name_map

    def apply(self, root_module):
        instances = []
        lookup = dict(root_module.named_modules())
        for linear_names, batch_norm_name in self.name_map:
            instance = self.copy()
            instance.register([lookup[name] for name in linear_names], lookup[batch_norm_name])
            instances.append(instance)
        return instances

    def copy(self):
        return self.__class__(self.name_map)


class CompositeCanonizer(Canonizer):
    def __init__(self, canonizers):
        self.canonizers = canonizers

    def apply(self, root_module):
        instances = []
        for canonizer in self.canonizers:
            instances += canonizer.apply(root_module)
        return instances
