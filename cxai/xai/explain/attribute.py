"""Attribution entry points -- mirror of the reference's cxai/xai/explain/attribute.py
(``SubspaceHook`` :12-67, ``compute_relevances`` :70-108, ``lrp_output_modifier`` :111-160)."""
from __future__ import annotations

import torch

__all__ = ["compute_relevances", "lrp_output_modifier", "SubspaceHook"]


class SubspaceHook:
    """Descriptor of the reference's backward hook on ``features.subspacefilter`` (attribute.py:12-67): the batch is
    read as groups of ``num_concepts + 1`` clones of one sample; clone 0 keeps all relevance, clone k only the
    relevance of concept block k of h.  The masking itself happens in ``lrp_subspace_filter`` (libdrsa_b200.so)."""
    kind = "subspace_hook"

    def __init__(self, num_concepts: int = 4, stabilizer: float = 1e-7, device=None) -> None:
        self.num_concepts = num_concepts
        self.stabilizer = stabilizer
        self.device = device

    def copy(self):
        return self.__class__(num_concepts=self.num_concepts, stabilizer=self.stabilizer, device=self.device)


def compute_relevances(model, input_batch: torch.Tensor, composite, num_classes: int = None, class_idx: int = None,
                       one_hot_encoded: bool = False) -> torch.Tensor:
    """LRP relevance of the selected output logit(s) at the input, same shape as ``input_batch``
    (attribute.py:70-108).  The batch holds samples of one class (``class_idx``) or equally many
    consecutive samples of every class (``num_classes``)."""
    from cxai.xai.explain.lrp_engine import lrp_input_relevance
    return lrp_input_relevance(model, input_batch, composite,
                               lrp_output_modifier(class_idx, num_classes, one_hot_encoded))


def lrp_output_modifier(class_idx: int = None, num_classes: int = None, one_hot_encoded: bool = False):
    """Returns the function that turns the logits into the relevance seed (attribute.py:111-160):
    the selected logit (or 1 if ``one_hot_encoded``) at the chosen class, zero elsewhere."""
    assert class_idx is not None or num_classes is not None, \
        "Provide either class_idx to attribute or num_classes to build the per-sample class mask"

    if class_idx is not None:
        def extract_output_class(output):
            mask = torch.zeros_like(output)
            mask[..., class_idx] = 1
            return mask if one_hot_encoded else output * mask
        extract_output_class.rowwise = True         # the seed of a row does not depend on the rest of the batch
        extract_output_class.key = ("class", int(class_idx), bool(one_hot_encoded))   # identifies the seed (graph replay)
        return extract_output_class

    def attribute_all_classes(output):
        mask = torch.repeat_interleave(torch.eye(num_classes).to(output), output.size(0) // num_classes, dim=0)
        return mask if one_hot_encoded else output * mask
    return attribute_all_classes
