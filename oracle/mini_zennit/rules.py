"""zennit.rules (0.5.1), restated.  TEST INFRASTRUCTURE ONLY.

Every rule is a ``BasicHook`` configuration: lists of input / parameter / output modifiers (one entry per modified forward
pass), a gradient mapper (relevance at the output -> gradient seeds of the passes) and a reducer (gradients -> relevance at
the input).  ``stabilizer`` / ``epsilon`` floats go through ``Stabilizer.ensure``."""
from __future__ import annotations

import torch

from .core import BasicHook, Hook, ParamMod, Stabilizer, zero_bias

# Could not be settled from memory of the 0.5.1 sources: does ``Flat`` set the bias to one (like every other parameter)
# or drop it?  False = bias set to one (what the CUDA path and oracle/lrp_ref.py do).  Only the input-layer relevance of
# the toy model (``LRP_NAME_MAP_TOY``, constants.py:42) depends on it, never the context vectors at a split layer.
FLAT_ZERO_BIAS = False


class NoMod(ParamMod):
    def __init__(self, zero_params=None, param_keys=None, require_params=False):
        super().__init__(lambda param, name: param, param_keys=param_keys, require_params=require_params,
                         zero_params=zero_params)


class ClampMod(ParamMod):
    def __init__(self, min=None, max=None, **kwargs):
        super().__init__(lambda param, name: param.clamp(min=min, max=max), **kwargs)


class GammaMod(ParamMod):
    def __init__(self, gamma=0.25, min=None, max=None, **kwargs):
        super().__init__(lambda param, name: param + gamma * param.clamp(min=min, max=max), **kwargs)


class Epsilon(BasicHook):
    """R_in = x * grad( R_out / stabilize(z) )."""

    def __init__(self, epsilon=1e-6, zero_params=None):
        stabilizer_fn = Stabilizer.ensure(epsilon)
        super().__init__(
            input_modifiers=[lambda input: input],
            param_modifiers=[NoMod(zero_params=zero_params)],
            output_modifiers=[lambda output: output],
            gradient_mapper=(lambda out_grad, outputs: out_grad / stabilizer_fn(outputs[0])),
            reducer=(lambda inputs, gradients: inputs[0] * gradients[0]),
        )


class Gamma(BasicHook):
    """Generalised gamma rule: four modified passes select the positive / negative contributions, a fifth plain pass
    decides the sign of the output."""

    def __init__(self, gamma=0.25, stabilizer=1e-6, zero_params=None):
        mod_kwargs = {'zero_params': zero_params}
        mod_kwargs_nobias = {'zero_params': zero_bias(zero_params)}
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        super().__init__(
            input_modifiers=[
                lambda input: input.clamp(min=0),
                lambda input: input.clamp(max=0),
                lambda input: input.clamp(min=0),
                lambda input: input.clamp(max=0),
                lambda input: input,
            ],
            param_modifiers=[
                GammaMod(gamma, min=0., **mod_kwargs),
                GammaMod(gamma, max=0., **mod_kwargs_nobias),
                GammaMod(gamma, max=0., **mod_kwargs),
                GammaMod(gamma, min=0., **mod_kwargs_nobias),
                NoMod(),
            ],
            output_modifiers=[lambda output: output] * 5,
            gradient_mapper=(
                lambda out_grad, outputs: [
                    output * out_grad / stabilizer_fn(denom)
                    for output, denom in (
                        [(outputs[4] > 0., sum(outputs[:2]))] * 2
                        + [(outputs[4] < 0., sum(outputs[2:4]))] * 2
                    )
                ] + [torch.zeros_like(out_grad)]
            ),
            reducer=(lambda inputs, gradients: sum(input * gradient
                                                   for input, gradient in zip(inputs[:4], gradients[:4]))),
        )


class ZPlus(BasicHook):
    def __init__(self, stabilizer=1e-6, zero_params=None):
        mod_kwargs = {'zero_params': zero_params}
        mod_kwargs_nobias = {'zero_params': zero_bias(zero_params)}
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        super().__init__(
            input_modifiers=[lambda input: input.clamp(min=0), lambda input: input.clamp(max=0)],
            param_modifiers=[ClampMod(min=0., **mod_kwargs), ClampMod(max=0., **mod_kwargs_nobias)],
            output_modifiers=[lambda output: output] * 2,
            gradient_mapper=(lambda out_grad, outputs: [out_grad / stabilizer_fn(sum(outputs))] * 2),
            reducer=(lambda inputs, gradients: inputs[0] * gradients[0] + inputs[1] * gradients[1]),
        )


class AlphaBeta(BasicHook):
    def __init__(self, alpha=2., beta=1., stabilizer=1e-6, zero_params=None):
        if alpha < 0 or beta < 0:
            raise ValueError("Both alpha and beta parameters must be non-negative!")
        if (alpha - beta) != 1.:
            raise ValueError("The difference of parameters alpha - beta must equal 1!")
        mod_kwargs = {'zero_params': zero_params}
        mod_kwargs_nobias = {'zero_params': zero_bias(zero_params)}
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        super().__init__(
            input_modifiers=[
                lambda input: input.clamp(min=0),
                lambda input: input.clamp(max=0),
                lambda input: input.clamp(min=0),
                lambda input: input.clamp(max=0),
            ],
            param_modifiers=[
                ClampMod(min=0., **mod_kwargs),
                ClampMod(max=0., **mod_kwargs_nobias),
                ClampMod(max=0., **mod_kwargs),
                ClampMod(min=0., **mod_kwargs_nobias),
            ],
            output_modifiers=[lambda output: output] * 4,
            gradient_mapper=(
                lambda out_grad, outputs: [
                    out_grad / stabilizer_fn(denom)
                    for denom in ([sum(outputs[:2])] * 2 + [sum(outputs[2:])] * 2)
                ]
            ),
            reducer=(
                lambda inputs, gradients: (
                    alpha * (inputs[0] * gradients[0] + inputs[1] * gradients[1])
                    - beta * (inputs[2] * gradients[2] + inputs[3] * gradients[3])
                )
            ),
        )


class ZBox(BasicHook):
    def __init__(self, low, high, stabilizer=1e-6, zero_params=None):
        def sub(positive, *negatives):
            return positive - sum(negatives)

        mod_kwargs = {'zero_params': zero_params}
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        super().__init__(
            input_modifiers=[
                lambda input: input,
                lambda input: low[:input.shape[0]] if isinstance(low, torch.Tensor) and low.dim() else
                torch.full_like(input, float(low)),
                lambda input: high[:input.shape[0]] if isinstance(high, torch.Tensor) and high.dim() else
                torch.full_like(input, float(high)),
            ],
            param_modifiers=[
                NoMod(**mod_kwargs),
                ClampMod(min=0., **mod_kwargs),
                ClampMod(max=0., **mod_kwargs),
            ],
            output_modifiers=[lambda output: output] * 3,
            gradient_mapper=(lambda out_grad, outputs: (out_grad / stabilizer_fn(sub(*outputs)),) * 3),
            reducer=(lambda inputs, gradients: sub(*(input * gradient for input, gradient in zip(inputs, gradients)))),
        )


class Pass(Hook):
    """Passes the incoming relevance on unchanged."""

    def backward(self, module, grad_input, grad_output):
        return grad_output


class Norm(BasicHook):
    def __init__(self, stabilizer=1e-6):
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        super().__init__(
            input_modifiers=[lambda input: input],
            param_modifiers=[NoMod()],
            output_modifiers=[lambda output: output],
            gradient_mapper=(lambda out_grad, outputs: out_grad / stabilizer_fn(outputs[0])),
            reducer=(lambda inputs, gradients: inputs[0] * gradients[0]),
        )


class WSquare(BasicHook):
    """Input replaced by ones, parameters squared, no input factor."""

    def __init__(self, stabilizer=1e-6, zero_params=None):
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        super().__init__(
            input_modifiers=[torch.ones_like],
            param_modifiers=[ParamMod((lambda param, name: param ** 2), zero_params=zero_params)],
            output_modifiers=[lambda output: output],
            gradient_mapper=(lambda out_grad, outputs: out_grad / stabilizer_fn(outputs[0])),
            reducer=(lambda inputs, gradients: gradients[0]),
        )


class Flat(BasicHook):
    """As WSquare with every parameter set to one."""

    def __init__(self, stabilizer=1e-6, zero_params=None):
        stabilizer_fn = Stabilizer.ensure(stabilizer)
        zp = zero_bias(zero_params) if FLAT_ZERO_BIAS else zero_params
        super().__init__(
            input_modifiers=[torch.ones_like],
            param_modifiers=[ParamMod((lambda param, name: torch.ones_like(param)), zero_params=zp,
                                      require_params=False)],
            output_modifiers=[lambda output: output],
            gradient_mapper=(lambda out_grad, outputs: out_grad / stabilizer_fn(outputs[0])),
            reducer=(lambda inputs, gradients: gradients[0]),
        )
