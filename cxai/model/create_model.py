"""The log-mel CNN that is explained -- mirror of the reference's cxai/model/create_model.py
(``VGGType`` :8-97, ``get_conv_block_layers`` :100, ``get_dense_block_layers`` :140, ``get_out_shape`` :174).

The torch modules only hold the parameters and the layer order; the forward and LRP passes of the hot
path run in ``cxai.xai.explain.lrp_engine`` on the CUDA library.  Unlike the reference (hard-coded
``x.view(-1, 2048)``, create_model.py:95) the flatten size is derived from the
configuration, so the d=256 / d=512 variants of BASELINE.json can be built.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn


class VGGType(nn.Module):
    def __init__(self, n_filters: List[int] = [32, 64, 96, 128], conv_kernel: Tuple[int, int] = (3, 3),
                 pool_kernels: List[Tuple[int, int]] = [(4, 4), (2, 4), (2, 2), (2, 2)], n_dense: int = 512,
                 n_classes: int = 10, dropout: float = 0.2, block_depth: int = 2, dense_depth: int = 2,
                 input_size: Tuple[int, int] = (128, 256), padding: str = "same", stride: int = 1,
                 conv_bn: bool = True, dense_bn: bool = True) -> None:
        super().__init__()
        assert len(n_filters) == len(pool_kernels), "one max-pool kernel per convolutional block"
        blocks = []
        for i, filters in enumerate(n_filters):
            blocks.extend(get_conv_block_layers(n_in=n_filters[i - 1] if i > 0 else 1, n_out=filters,
                                                kernel=conv_kernel, block_depth=block_depth, padding=padding,
                                                stride=stride, conv_bn=conv_bn))
            blocks.append(nn.MaxPool2d(pool_kernels[i]))
        self.features = nn.Sequential(*blocks)
        self.num_flat_features = get_out_shape(input_size=input_size, conv_kernel=conv_kernel,
                                               pool_kernels=pool_kernels, padding=padding,
                                               out_filters=n_filters[-1], block_depth=block_depth)
        self.classifier = nn.Sequential(
            *get_dense_block_layers(n_in=self.num_flat_features, n_out=n_dense, dropout=dropout, depth=dense_depth,
                                    dense_bn=dense_bn),
            nn.Linear(n_dense, n_classes))

    def forward(self, x):
        x = self.features(x)
        return self.classifier(x.reshape(x.size(0), -1))


def get_conv_block_layers(n_in: int, n_out: int, block_depth: int = 2, kernel: Tuple[int, int] = (3, 3),
                          stride: int = 1, padding=1, padding_mode: str = "zeros", conv_bn: bool = True):
    """[Conv2d -> (BatchNorm2d) -> ReLU] x block_depth."""
    layers = []
    for i in range(block_depth):
        layers.append(nn.Conv2d(n_in if i == 0 else n_out, n_out, kernel_size=kernel, stride=stride,
                                padding=padding, padding_mode=padding_mode))
        if conv_bn:
            layers.append(nn.BatchNorm2d(n_out))
        layers.append(nn.ReLU())
    return layers


def get_dense_block_layers(n_in: int, n_out: int, dropout: float, depth: int = 2, dense_bn: bool = True):
    """[Linear -> (BatchNorm1d) -> ReLU -> (Dropout)] x depth."""
    layers = []
    for i in range(depth):
        layers.append(nn.Linear(n_in if i == 0 else n_out, n_out))
        if dense_bn:
            layers.append(nn.BatchNorm1d(n_out))
        layers.append(nn.ReLU())
        if dropout:
            layers.append(nn.Dropout(dropout))
    return layers


def get_out_shape(input_size=(128, 216), conv_kernel=(3, 3), pool_kernels=[(4, 4), (2, 4), (2, 2), (2, 2)],
                  out_filters: int = 128, padding=1, stride: int = 1, block_depth: int = 2) -> int:
    """Number of features after the last pooling layer."""
    pad = 1 if padding == "same" else (padding if isinstance(padding, int) else 0)
    h, w = input_size
    for pk in pool_kernels:
        for _ in range(block_depth):
            h = (h - conv_kernel[0] + 2 * pad) // stride + 1
            w = (w - conv_kernel[1] + 2 * pad) // stride + 1
        h, w = h // pk[0], w // pk[1]
    return int(h * w * out_filters)


def num_flat_features(x: torch.Tensor) -> int:
    n = 1
    for s in x.size()[1:]:
        n *= s
    return n
