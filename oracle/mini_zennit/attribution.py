"""zennit.attribution (0.5.1), restated: the Gradient attributor.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch


def identity(obj):
    return obj


def constant(obj):
    def wrapped_const(*args, **kwargs):
        return obj
    return wrapped_const


class Attributor:
    def __init__(self, model, composite=None, attr_output=None):
        self.model = model
        self.composite = composite
        if attr_output is None:
            self.attr_output_fn = identity
        elif not callable(attr_output):
            self.attr_output_fn = constant(attr_output)
        else:
            self.attr_output_fn = attr_output

    def __enter__(self):
        if self.composite is not None:
            self.composite.register(self.model)
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        if self.composite is not None:
            self.composite.remove()
        return False

    def __call__(self, input, attr_output=None):
        if attr_output is None:
            attr_output_fn = self.attr_output_fn
        elif not callable(attr_output):
            attr_output_fn = constant(attr_output)
        else:
            attr_output_fn = attr_output
        if self.composite is not None and not self.composite.handles:
            with self:
                return self.forward(input, attr_output_fn)
        return self.forward(input, attr_output_fn)

    @property
    def inactive(self):
        return self.composite.inactive()

    def forward(self, input, attr_output_fn):
        raise NotImplementedError


class Gradient(Attributor):
    """out = model(input); attribution = d out / d input seeded with ``attr_output_fn(out)``."""

    def __init__(self, model, composite=None, attr_output=None, create_graph=False, retain_graph=None):
        super().__init__(model=model, composite=composite, attr_output=attr_output)
        self.create_graph = create_graph
        self.retain_graph = retain_graph

    def forward(self, input, attr_output_fn):
        if not input.requires_grad:
            input.requires_grad = True
        output = self.model(input)
        gradient, = torch.autograd.grad((output,), (input,), grad_outputs=(attr_output_fn(output.detach()),),
                                        create_graph=self.create_graph, retain_graph=self.retain_graph)
        return output, gradient
