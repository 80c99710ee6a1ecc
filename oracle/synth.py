"""Seeded synthetic inputs shared by the fixture generators and the tests.  TEST INFRASTRUCTURE ONLY.
Imports nothing from ``cxai`` (neither the product's nor the reference's), so it can be used in either process."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def synth_logmel(N: int, H: int, W: int, seed: int) -> torch.Tensor:
    """x = clamp(1.2*randn - 1.5, min=-4), [N,1,H,W] (value range of utils/dataloading.py:159-161)."""
    g = torch.Generator().manual_seed(seed)
    return (1.2 * torch.randn(N, 1, H, W, generator=g) - 1.5).clamp(min=-4.0)


def randomize_bn(net: nn.Module, seed: int) -> nn.Module:
    """Non-trivial BatchNorm statistics / affine parameters so that the BN fold is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            m.weight.data.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    return net


def weight_checksum(net: nn.Module) -> np.ndarray:
    """[sum, sum of |.|, sum of squares] over every parameter and buffer, in fp64: proves that two processes built
    the same model from the same seed."""
    acc = np.zeros(3)
    for _, t in sorted(list(net.state_dict().items())):
        t = t.double()
        acc += np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])
    return acc


def signed_permutation(d: int, seed: int) -> torch.Tensor:
    """Orthogonal matrix whose projections are exact in floating point."""
    g = torch.Generator().manual_seed(seed)
    P = torch.zeros(d, d)
    P[torch.arange(d), torch.randperm(d, generator=g)] = torch.where(torch.rand(d, generator=g) < 0.5, -1.0, 1.0)
    return P


def random_orthogonal(d: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.linalg.qr(torch.randn(d, d, generator=g, dtype=torch.float64))[0].float().contiguous()


# model configurations (constructor arguments of VGGType, create_model.py:8-97) used by fixtures and tests
MODEL_CONFIGS = {
    # cfg 1: 5 x [Conv3x3 - ReLU - MaxPool2] on 64x64, d = 64 at features[13], layer indices of LRP_NAME_MAP_TOY
    "toy": dict(n_filters=[8, 8, 16, 16, 64], pool_kernels=[(2, 2)] * 5, n_dense=64, n_classes=2, dropout=0.0,
                block_depth=1, dense_depth=2, input_size=(64, 64), conv_bn=False, dense_bn=False),
    # the reference's own toy geometry (flat size 64 = modify_model.py:38), used for the projection model
    "toy16": dict(n_filters=[8, 8, 16, 16, 16], pool_kernels=[(2, 2)] * 5, n_dense=64, n_classes=2, dropout=0.0,
                  block_depth=1, dense_depth=2, input_size=(64, 64), conv_bn=False, dense_bn=False),
    # arch A at reduced resolution (getdrsadata.py:72-73 with a 32x64 input and d = 64)
    "archA_small": dict(n_filters=[64, 64, 100, 128, 64], pool_kernels=[(2, 4), (2, 2), (2, 2), (2, 2), (2, 2)],
                        n_dense=100, n_classes=10, dropout=0.3, block_depth=2, dense_depth=2, input_size=(32, 64),
                        conv_bn=True, dense_bn=True),
    # the production model itself (getdrsadata.py:72-73): flat size 2048, d = 128 at features[33]
    "archA": dict(n_filters=[64, 64, 100, 128, 128], pool_kernels=[(2, 4), (2, 2), (2, 2), (2, 2), (2, 2)],
                  n_dense=100, n_classes=10, dropout=0.3, block_depth=2, dense_depth=2, input_size=(128, 256),
                  conv_bn=True, dense_bn=True),
    # arch B, the reference's 3-second GTZAN model (pixelflipping/cpf.py:410-412; rules LRP_NAME_MAP_GTZAN, constants.py:27-38):
    # one conv per block, no BatchNorm, 128 x 128 input, flat size 2048; split layers 1, 4, 7, 10, 13 (cpf.py:141)
    "archB": dict(n_filters=[32, 32, 64, 64, 128], pool_kernels=[(2, 2)] * 5, n_dense=128, n_classes=10, dropout=0.4,
                  block_depth=1, dense_depth=2, input_size=(128, 128), conv_bn=False, dense_bn=False),
    # BASELINE cfg 2: arch A with the last block widened to d = 256
    "cfg2": dict(n_filters=[64, 64, 100, 128, 256], pool_kernels=[(2, 4), (2, 2), (2, 2), (2, 2), (2, 2)],
                 n_dense=100, n_classes=10, dropout=0.3, block_depth=2, dense_depth=2, input_size=(128, 256),
                 conv_bn=True, dense_bn=True),
}


def build_model(vgg_cls, name: str, seed: int = 0, bn_seed: int | None = 1):
    """``vgg_cls(**MODEL_CONFIGS[name])`` under ``torch.manual_seed(seed)``, BatchNorm randomised, eval mode."""
    torch.manual_seed(seed)
    net = vgg_cls(**MODEL_CONFIGS[name])
    if bn_seed is not None:
        randomize_bn(net, bn_seed)
    return net.eval()
