"""Attribution entry points -- mirror of the reference's cxai/xai/explain/attribute.py
(``compute_relevances`` :70-108, ``lrp_output_modifier`` :111-160).  ``SubspaceHook`` (:12-67) belongs to
the concept-heatmap consumer, a "next" row of SURVEY section 8(f)."""
from __future__ import annotations

import torch

__all__ = ["compute_relevances", "lrp_output_modifier"]


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
        return extract_output_class

    def attribute_all_classes(output):
        mask = torch.repeat_interleave(torch.eye(num_classes).to(output), output.size(0) // num_classes, dim=0)
        return mask if one_hot_encoded else output * mask
    return attribute_all_classes
