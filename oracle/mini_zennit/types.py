"""zennit.types, restated: abstract module groups matched with isinstance.  TEST INFRASTRUCTURE ONLY."""
import torch


class SubclassMeta(type):
    def __instancecheck__(cls, inst):
        return cls.__subclasscheck__(type(inst))

    def __subclasscheck__(cls, sub):
        candidates = cls.__dict__.get("__subclass__", tuple())
        return type.__subclasscheck__(cls, sub) or issubclass(sub, candidates)


class ConvolutionTranspose(metaclass=SubclassMeta):
    __subclass__ = (torch.nn.modules.conv.ConvTranspose1d, torch.nn.modules.conv.ConvTranspose2d,
                    torch.nn.modules.conv.ConvTranspose3d)


class ConvolutionStandard(metaclass=SubclassMeta):
    __subclass__ = (torch.nn.modules.conv.Conv1d, torch.nn.modules.conv.Conv2d, torch.nn.modules.conv.Conv3d)


class Convolution(metaclass=SubclassMeta):
    __subclass__ = (ConvolutionStandard, ConvolutionTranspose)


class Linear(metaclass=SubclassMeta):
    __subclass__ = (Convolution, torch.nn.modules.linear.Linear)


class BatchNorm(metaclass=SubclassMeta):
    __subclass__ = (torch.nn.modules.batchnorm.BatchNorm1d, torch.nn.modules.batchnorm.BatchNorm2d,
                    torch.nn.modules.batchnorm.BatchNorm3d)


class AvgPool(metaclass=SubclassMeta):
    __subclass__ = (torch.nn.modules.pooling.AvgPool1d, torch.nn.modules.pooling.AvgPool2d,
                    torch.nn.modules.pooling.AvgPool3d, torch.nn.modules.pooling.AdaptiveAvgPool1d,
                    torch.nn.modules.pooling.AdaptiveAvgPool2d, torch.nn.modules.pooling.AdaptiveAvgPool3d)


class MaxPool(metaclass=SubclassMeta):
    __subclass__ = (torch.nn.modules.pooling.MaxPool1d, torch.nn.modules.pooling.MaxPool2d,
                    torch.nn.modules.pooling.MaxPool3d)


class Activation(metaclass=SubclassMeta):
    __subclass__ = (torch.nn.modules.activation.ELU, torch.nn.modules.activation.Hardtanh,
                    torch.nn.modules.activation.LeakyReLU, torch.nn.modules.activation.ReLU,
                    torch.nn.modules.activation.ReLU6, torch.nn.modules.activation.Sigmoid,
                    torch.nn.modules.activation.Tanh, torch.nn.modules.activation.Softplus,
                    torch.nn.modules.activation.GELU, torch.nn.modules.activation.SiLU)
