"""zennit.composites (0.5.1), restated.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

from .core import Composite  # noqa: F401  (re-exported: ``from zennit.composites import Composite``)


class LayerMapComposite(Composite):
    """Rules by module type: ``layer_map = [(types, hook), ...]``, first match wins."""

    def __init__(self, layer_map, canonizers=None):
        self.layer_map = layer_map
        super().__init__(module_map=self.mapping, canonizers=canonizers)

    def mapping(self, ctx, name, module):
        return next((hook for types, hook in self.layer_map if isinstance(module, types)), None)


class SpecialFirstLayerMapComposite(LayerMapComposite):
    """As LayerMapComposite with a separate map for the first module that matches it."""

    def __init__(self, layer_map, first_map, canonizers=None):
        self.first_map = first_map
        super().__init__(layer_map=layer_map, canonizers=canonizers)

    def mapping(self, ctx, name, module):
        if not ctx.get('first_layer_visited', False):
            for types, hook in self.first_map:
                if isinstance(module, types):
                    ctx['first_layer_visited'] = True
                    return hook
        return super().mapping(ctx, name, module)


class NameMapComposite(Composite):
    """Rules by module name: ``name_map = [(names, hook), ...]``."""

    def __init__(self, name_map, canonizers=None):
        self.name_map = name_map
        super().__init__(module_map=self.mapping, canonizers=canonizers)

    def mapping(self, ctx, name, module):
        return next((hook for names, hook in self.name_map if name in names), None)


class NameLayerMapComposite(Composite):
    """Name map first, type map as the fall-back."""

    def __init__(self, name_map=None, layer_map=None, canonizers=None):
        self.name_map = name_map if name_map is not None else []
        self.layer_map = layer_map if layer_map is not None else []
        super().__init__(module_map=self.mapping, canonizers=canonizers)

    def mapping(self, ctx, name, module):
        hook = next((hook for names, hook in self.name_map if name in names), None)
        if hook is None:
            hook = next((hook for types, hook in self.layer_map if isinstance(module, types)), None)
        return hook
