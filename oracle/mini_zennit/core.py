"""zennit.core (0.5.1), restated: stabilize, Identity, Hook, BasicHook, ParamMod, Composite.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import functools
import weakref
from contextlib import contextmanager

import torch


def stabilize(input, epsilon=1e-6, clip=False, norm_scale=False, dim=None):
    """x + ((x == 0) + sign(x)) * eps: zero counts as positive."""
    sign = ((input == 0.).to(input) + input.sign())
    if norm_scale:
        if dim is None:
            dim = tuple(range(1, input.ndim))
        epsilon = epsilon * ((input ** 2).mean(dim=dim, keepdim=True) ** .5)
    if clip:
        return sign * input.abs().clip(min=epsilon)
    return input + sign * epsilon


class Stabilizer:
    def __init__(self, epsilon=1e-6, clip=False, norm_scale=False, dim=None):
        self.epsilon, self.clip, self.norm_scale, self.dim = epsilon, clip, norm_scale, dim

    def __call__(self, input):
        return stabilize(input, self.epsilon, self.clip, self.norm_scale, self.dim)

    @classmethod
    def ensure(cls, value):
        if isinstance(value, (float, int)):
            return cls(epsilon=float(value))
        if callable(value):
            return value
        raise TypeError(f"Value {value} is not a valid stabilizer!")


class Identity(torch.autograd.Function):
    """Identity that guarantees a grad_fn on which the backward hook of a rule can be registered."""

    @staticmethod
    def forward(ctx, *inputs):
        return inputs

    @staticmethod
    def backward(ctx, *grad_outputs):
        return grad_outputs


class RemovableHandle:
    def __init__(self, instance):
        self.instance_ref = weakref.ref(instance)

    def remove(self):
        instance = self.instance_ref()
        if instance is not None:
            instance.remove()


class RemovableHandleList(list):
    def remove(self):
        for handle in self:
            handle.remove()
        self.clear()


class Hook:
    """Base hook: wraps the module input in an identity node, and replaces the gradient w.r.t. the module input by
    ``self.backward(module, grad_input, grad_output)`` during the backward pass."""

    def __init__(self):
        self.stored_tensors = {}
        self.active = True
        self.tensor_handles = RemovableHandleList()

    def pre_forward(self, module, input):
        hook_ref = weakref.ref(self)

        @functools.wraps(self.backward)
        def wrapper(grad_input, grad_output):
            hook = hook_ref()
            if hook is not None and hook.active:
                return hook.backward(module, grad_input, hook.stored_tensors['grad_output'])
            return None

        if not isinstance(input, tuple):
            input = (input,)
        if input[0].requires_grad:                    # only if a gradient is required
            post_input = Identity.apply(*input)
            self.tensor_handles.append(post_input[0].grad_fn.register_hook(wrapper))
            post_input = tuple(elem.clone() for elem in post_input)      # supports in-place modules
        else:
            post_input = input
        return post_input[0] if len(post_input) == 1 else post_input

    def post_forward(self, module, input, output):
        hook_ref = weakref.ref(self)

        @functools.wraps(self.pre_backward)
        def wrapper(grad_input, grad_output):
            hook = hook_ref()
            if hook is not None and hook.active:
                return hook.pre_backward(module, grad_input, grad_output)
            return None

        if not isinstance(output, tuple):
            output = (output,)
        if output[0].grad_fn is not None:
            self.tensor_handles.append(output[0].grad_fn.register_hook(wrapper))
        return output[0] if len(output) == 1 else output

    def pre_backward(self, module, grad_input, grad_output):
        self.stored_tensors['grad_output'] = grad_output

    def forward(self, module, input, output):
        """hook applied during the forward pass"""

    def backward(self, module, grad_input, grad_output):
        """hook applied during the backward pass"""

    def copy(self):
        return self.__class__()

    def remove(self):
        self.tensor_handles.remove()

    def register(self, module):
        return RemovableHandleList([
            RemovableHandle(self),
            module.register_forward_pre_hook(self.pre_forward),
            module.register_forward_hook(self.post_forward),
            module.register_forward_hook(self.forward),
        ])


def zero_wrap(zero_params):
    """Decorator factory: parameters whose name is listed come back as zeros."""
    if zero_params is None:
        zero_params = []
    elif isinstance(zero_params, str):
        zero_params = [zero_params]

    def wrapper(modifier):
        @functools.wraps(modifier)
        def modifier_wrapper(input, name):
            if name in zero_params:
                return torch.zeros_like(input)
            return modifier(input, name)
        return modifier_wrapper
    return wrapper


def zero_bias(zero_params=None):
    """``zero_params`` with 'bias' added."""
    if zero_params is None:
        return ['bias']
    if isinstance(zero_params, str):
        zero_params = [zero_params]
    return list(set(list(zero_params) + ['bias']))


class ParamMod:
    """Context manager that temporarily replaces the parameters of a module by ``modifier(param, name)``."""

    def __init__(self, modifier, param_keys=None, require_params=True, zero_params=None):
        self.modifier = zero_wrap(zero_params)(modifier)
        self.param_keys = param_keys
        self.require_params = require_params

    @classmethod
    def ensure(cls, modifier):
        if isinstance(modifier, cls):
            return modifier
        if callable(modifier):
            return cls(modifier)
        raise TypeError(f"{modifier} is neither a ParamMod nor callable")

    @contextmanager
    def __call__(self, module):
        stored = {}
        try:
            param_keys = self.param_keys
            if param_keys is None:
                param_keys = [name for name, _ in module.named_parameters(recurse=False)]
            missing = [key for key in param_keys if not hasattr(module, key)]
            if self.require_params and missing:
                raise RuntimeError(f"Module {module} requires missing parameters: {missing}")
            for key in param_keys:
                if key in missing:
                    continue
                param = getattr(module, key)
                if param is not None:
                    stored[key] = param
                    # shadow the registered parameter through the instance dict (found before nn.Module.__getattr__)
                    object.__setattr__(module, key, self.modifier(param.data, key))
            yield module
        finally:
            for key in stored:
                object.__delattr__(module, key)


def collect_leaves(module):
    """Leaf modules of ``module`` in registration order."""
    is_leaf = True
    for child in module.children():
        is_leaf = False
        yield from collect_leaves(child)
    if is_leaf:
        yield module


class BasicHook(Hook):
    """Rule skeleton: for every (input modifier, parameter modifier, output modifier) triple run the module forward on
    modified inputs / parameters under autograd, map the incoming relevance to gradient seeds, take the gradients and
    reduce them to the relevance at the input."""

    def __init__(self, input_modifiers=None, param_modifiers=None, output_modifiers=None, gradient_mapper=None,
                 reducer=None):
        super().__init__()
        modifiers = {'in': input_modifiers, 'param': param_modifiers, 'out': output_modifiers}
        supplied = {key for key, val in modifiers.items() if val is not None}
        num_mods = len(modifiers[next(iter(supplied))]) if supplied else 1
        modifiers.update({key: (self._default_modifier,) * num_mods for key in set(modifiers) - supplied})
        self.input_modifiers = modifiers['in']
        self.param_modifiers = modifiers['param']
        self.output_modifiers = modifiers['out']
        self.gradient_mapper = gradient_mapper if gradient_mapper is not None else self._default_gradient_mapper
        self.reducer = reducer if reducer is not None else self._default_reducer

    def forward(self, module, input, output):
        self.stored_tensors['input'] = input

    def backward(self, module, grad_input, grad_output):
        original_input = self.stored_tensors['input'][0].clone()
        inputs, outputs = [], []
        for in_mod, param_mod, out_mod in zip(self.input_modifiers, self.param_modifiers, self.output_modifiers):
            input = in_mod(original_input).requires_grad_()
            with ParamMod.ensure(param_mod)(module) as modified, torch.autograd.enable_grad():
                output = modified.forward(input)
                output = out_mod(output)
            inputs.append(input)
            outputs.append(output)
        grad_outputs = self.gradient_mapper(grad_output[0], outputs)
        gradients = torch.autograd.grad(outputs, inputs, grad_outputs=grad_outputs,
                                        create_graph=grad_output[0].requires_grad)
        relevance = self.reducer(inputs, gradients)
        return tuple(relevance if original.shape == relevance.shape else None for original in grad_input)

    def copy(self):
        copy = BasicHook.__new__(type(self))
        BasicHook.__init__(copy, self.input_modifiers, self.param_modifiers, self.output_modifiers, self.gradient_mapper,
                           self.reducer)
        return copy

    @staticmethod
    def _default_modifier(obj, name=None):
        return obj

    @staticmethod
    def _default_gradient_mapper(out_grad, outputs):
        return tuple(out_grad / stabilize(output) for output in outputs)

    @staticmethod
    def _default_reducer(inputs, gradients):
        return sum(input * gradient for input, gradient in zip(inputs, gradients))


class Composite:
    """Maps hooks to the modules of a model (``module_map(ctx, name, module) -> hook template or None``) and applies
    canonizers first."""

    def __init__(self, module_map=None, canonizers=None):
        self.module_map = module_map if module_map is not None else (lambda ctx, name, module: None)
        self.canonizers = canonizers if canonizers is not None else []
        self.handles = RemovableHandleList()
        self.hook_refs = weakref.WeakSet()

    def register(self, module):
        self.remove()
        for canonizer in self.canonizers:
            self.handles += canonizer.apply(module)
        ctx = {}
        for name, child in module.named_modules():
            template = self.module_map(ctx, name, child)
            if template is not None:
                hook = template.copy()
                self.hook_refs.add(hook)
                self.handles.append(hook.register(child))

    def remove(self):
        self.handles.remove()
        self.hook_refs.clear()

    def context(self, module):
        return CompositeContext(module, self)

    @contextmanager
    def inactive(self):
        try:
            for hook in self.hook_refs:
                hook.active = False
            yield self
        finally:
            for hook in self.hook_refs:
                hook.active = True


class CompositeContext:
    def __init__(self, module, composite):
        self.module, self.composite = module, composite

    def __enter__(self):
        self.composite.register(self.module)
        return self.module

    def __exit__(self, exc_type, exc_value, traceback):
        self.composite.remove()
        return False
