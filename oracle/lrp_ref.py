"""CPU restatement of the LRP pass the reference runs through zennit 0.5.1 (stage 1 of the hot path).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED: all LRP arithmetic of the reference lives in the third-party package
``zennit==0.5.1`` (reference ``requirements.txt:23``), which is neither vendored under
/root/reference nor installed nor obtainable here, and the reference has no tests or golden
vectors for this path.  This file restates zennit 0.5.1's published rule
semantics and is anchored on the reference's own call sites:

  * orchestration  -- ``get_intermediate`` (cxai/xai/drsa/preprocessing.py:106-176): forward hook on
    ``layer`` keeps ``layer.output`` and its ``.grad``; minibatches of 64; the relevance seed comes from
    ``lrp_output_modifier`` (cxai/xai/explain/attribute.py:111-160);
  * rule <-> layer maps -- cxai/utils/constants.py:27-51, cxai/xai/drsa/cluster/getdrsadata.py:87-108;
  * model -- cxai/model/create_model.py:100-171.

Unlike the CUDA path (collapsed one-pass Gamma for non-negative inputs) this oracle implements the
GENERAL multi-pass form of every rule exactly as zennit's ``BasicHook.backward`` does it: for each
(input modifier, parameter modifier) pair run the module forward on modified inputs/parameters under
autograd, map the incoming relevance to gradient seeds, call ``torch.autograd.grad`` and reduce.
Agreement of the two is therefore a real check of the collapse argument.
Self-checks (tests/test_oracle_lrp.py): relevance conservation, collapsed == general for x >= 0,
Epsilon == gradient*input for eps -> 0 on a bias-free ReLU net.
"""
from __future__ import annotations

import copy
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def stabilize(x: torch.Tensor, eps: float) -> torch.Tensor:
    """zennit.core.stabilize: x + ((x == 0) + sign(x)) * eps -- zero counts as positive."""
    return x + ((x == 0.).to(x) + x.sign()) * eps


# --------------------------------------------------------------------------- rule table
# each rule: list of (input_modifier, weight_modifier, bias_modifier or None=drop bias),
#            gradient mapper (R_out, outputs) -> list of grad_outputs, reducer (inputs, grads) -> R_in
def _rule_passes(kind: str, gamma: float = 0.0):
    pos = lambda t: t.clamp(min=0)
    neg = lambda t: t.clamp(max=0)
    ident = lambda t: t
    if kind == "epsilon":
        return [(ident, ident, ident)]
    if kind == "gamma":       # zennit Gamma (generalised): 4 modified passes + 1 plain pass
        return [(pos, lambda w: w + gamma * pos(w), lambda b: b + gamma * pos(b)),
                (neg, lambda w: w + gamma * neg(w), None),
                (pos, lambda w: w + gamma * neg(w), lambda b: b + gamma * neg(b)),
                (neg, lambda w: w + gamma * pos(w), None),
                (ident, ident, ident)]
    if kind == "zplus":
        return [(pos, pos, pos), (neg, neg, None)]
    if kind == "wsquare":
        return [(torch.ones_like, lambda w: w ** 2, lambda b: b ** 2)]
    if kind == "flat":
        return [(torch.ones_like, torch.ones_like, torch.ones_like)]
    raise ValueError(kind)


def _module_forward(module: nn.Module, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]):
    if isinstance(module, nn.Conv2d):
        pad = module.padding if not isinstance(module.padding, str) else module.padding
        return F.conv2d(x, w, b, stride=module.stride, padding=pad, dilation=module.dilation, groups=module.groups)
    if isinstance(module, nn.Linear):
        return F.linear(x, w, b)
    raise TypeError(type(module))


def rule_backward(module: nn.Module, rule_kind: str, x: torch.Tensor, R_out: torch.Tensor, eps: float,
                  gamma: float = 0.0, weight=None, bias=None) -> torch.Tensor:
    """BasicHook.backward of zennit for one layer (general multi-pass form)."""
    w = module.weight.detach() if weight is None else weight
    b = (module.bias.detach() if module.bias is not None else None) if bias is None and weight is None else bias
    passes = _rule_passes(rule_kind, gamma)
    inputs, outputs = [], []
    with torch.enable_grad():
        for in_mod, w_mod, b_mod in passes:
            xi = in_mod(x.detach()).requires_grad_(True)
            bi = b_mod(b) if (b is not None and b_mod is not None) else None
            inputs.append(xi)
            outputs.append(_module_forward(module, xi, w_mod(w), bi))
        if rule_kind == "gamma":
            z = outputs[4]
            gp = (z > 0.).to(z) * R_out / stabilize(outputs[0] + outputs[1], eps)
            gn = (z < 0.).to(z) * R_out / stabilize(outputs[2] + outputs[3], eps)
            grad_outputs = [gp, gp, gn, gn, torch.zeros_like(z)]
        elif rule_kind == "zplus":
            g = R_out / stabilize(outputs[0] + outputs[1], eps)
            grad_outputs = [g, g]
        else:
            grad_outputs = [R_out / stabilize(outputs[0], eps)]
        grads = torch.autograd.grad(outputs, inputs, grad_outputs=grad_outputs)
    if rule_kind in ("wsquare", "flat"):
        return grads[0]                                   # reducer: gradient only (no input factor)
    n_used = 4 if rule_kind == "gamma" else len(grads)
    return sum(inputs[i].detach() * grads[i] for i in range(n_used))


# --------------------------------------------------------------------------- network pass
def merge_batchnorm(layers: Sequence[Tuple[str, Optional[nn.Module]]]):
    """SequentialMergeBatchNorm: returns {name: (weight, bias)} for every Conv/Linear directly followed by
    a BatchNorm (eval mode) and the set of BN names that become identity."""
    merged, identity = {}, set()
    for (n0, m0), (n1, m1) in zip(layers[:-1], layers[1:]):
        if isinstance(m0, (nn.Conv2d, nn.Linear)) and isinstance(m1, (nn.BatchNorm1d, nn.BatchNorm2d)):
            scale = m1.weight.detach() / torch.sqrt(m1.running_var + m1.eps)
            w = m0.weight.detach() * scale.view(-1, *([1] * (m0.weight.dim() - 1)))
            b0 = m0.bias.detach() if m0.bias is not None else torch.zeros_like(m1.running_mean)
            b = (b0 - m1.running_mean) * scale + m1.bias.detach()
            merged[n0] = (w, b)
            identity.add(n1)
    return merged, identity


def lrp_pass(model: nn.Module, x: torch.Tensor, name_map, seed_fn: Callable, split_module: Optional[nn.Module] = None,
             merge_bn: bool = True, dtype=torch.float64):
    """One LRP pass of a VGGType-like model (``model.features`` + flatten + ``model.classifier``).

    name_map: [(names, rule)] with rule objects exposing .kind/.stabilizer/.gamma (cxai.xai.explain.rules).
    Returns dict(logits, R_input, a_split, R_split)."""
    rules = {}
    for names, rule in name_map:
        for n in names:
            rules[n] = rule
    layers = [(f"features.{n}", m) for n, m in model.features.named_children()] + [("flatten", None)] + \
             [(f"classifier.{n}", m) for n, m in model.classifier.named_children()]
    merged, identity = merge_batchnorm(layers) if merge_bn else ({}, set())
    cast = lambda t: t.to(dtype)
    # forward, remembering every layer input
    acts = [cast(x)]
    cur = acts[0]
    flat_shape = None
    for name, m in layers:
        if m is None:
            flat_shape = cur.shape
            cur = cur.reshape(cur.size(0), -1)
        elif isinstance(m, (nn.Conv2d, nn.Linear)):
            w, b = merged.get(name, (m.weight.detach(), m.bias.detach() if m.bias is not None else None))
            cur = _module_forward(m, cur, cast(w), cast(b) if b is not None else None)
        elif name in identity or isinstance(m, nn.Dropout):
            pass
        elif isinstance(m, nn.ReLU):
            cur = cur.clamp(min=0)
        elif isinstance(m, nn.MaxPool2d):
            cur = F.max_pool2d(cur, m.kernel_size, m.stride, m.padding)
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            cur = F.batch_norm(cur, cast(m.running_mean), cast(m.running_var), cast(m.weight), cast(m.bias), False, 0.0, m.eps)
        elif type(m).__name__ in ("Projection", "SubspaceFilter", "InvProjection"):
            cur = m(cur)            # modify_model.py:62-123 (U follows the dtype of the activations)
        else:
            raise TypeError(type(m))
        acts.append(cur)
    logits = cur
    R = seed_fn(logits.detach())
    out = {"logits": logits, "a_split": None, "R_split": None}
    for k in range(len(layers) - 1, -1, -1):
        name, m = layers[k]
        if split_module is not None and m is split_module:
            out["a_split"], out["R_split"] = acts[k + 1], R       # output of the layer and its .grad
        xin = acts[k]
        if m is None:
            R = R.reshape(flat_shape)
        elif isinstance(m, (nn.Conv2d, nn.Linear)):
            rule = rules.get(name)
            if rule is None:
                raise KeyError(f"no rule for {name}")
            if rule.kind == "pass":
                continue
            w, b = merged.get(name, (m.weight.detach(), m.bias.detach() if m.bias is not None else None))
            R = rule_backward(m, rule.kind, xin, R, rule.stabilizer, getattr(rule, "gamma", 0.0), cast(w),
                              cast(b) if b is not None else None)
        elif name in identity or isinstance(m, nn.Dropout):
            pass
        elif isinstance(m, nn.ReLU):
            R = R * (acts[k + 1] > 0).to(R)
        elif isinstance(m, nn.MaxPool2d):
            with torch.enable_grad():
                xi = xin.detach().requires_grad_(True)
                yo = F.max_pool2d(xi, m.kernel_size, m.stride, m.padding)
                R, = torch.autograd.grad(yo, xi, R)
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            scale = cast(m.weight) / torch.sqrt(cast(m.running_var) + m.eps)
            R = R * scale.view(1, -1, *([1] * (R.dim() - 2)))
        elif type(m).__name__ in ("Projection", "InvProjection"):
            # zennit Epsilon (BasicHook) on a parameter-free module: R_in = x * d/dx [module(x)] . (R_out / stabilize(module(x)))
            rule = rules[name]
            assert rule.kind == "epsilon"
            with torch.enable_grad():
                xi = xin.detach().requires_grad_(True)
                yo = m(xi)
                grad, = torch.autograd.grad(yo, xi, R / stabilize(yo.detach(), rule.stabilizer))
            R = xi.detach() * grad
        elif type(m).__name__ == "SubspaceFilter":
            # SubspaceHook.backward, attribute.py:42-60, verbatim
            K = rules[name].num_concepts
            batch, num_vecs, c, d_c = R.size()
            R = R.clone().view(-1, K + 1, num_vecs, c, d_c)
            R[:, 1:] *= torch.eye(K, dtype=R.dtype)[None, :, None, :, None]
            R = R.view(batch, num_vecs, c, d_c)
    out["R_input"] = R
    return out


def get_intermediate(model, input_batch, name_map, layer, class_idx, attr_batch_size: int = 64,
                     one_hot_encoded: bool = False, dtype=torch.float64):
    """Restatement of preprocessing.py:106-176 on top of ``lrp_pass``."""
    a_maps, r_maps = [], []
    seed = output_modifier(class_idx, None, one_hot_encoded)
    for i in range(0, input_batch.size(0), attr_batch_size):
        o = lrp_pass(model, input_batch[i:i + attr_batch_size], name_map, seed, split_module=layer, dtype=dtype)
        a_maps.append(o["a_split"]); r_maps.append(o["R_split"])
    return torch.cat(a_maps, 0), torch.cat(r_maps, 0)


def output_modifier(class_idx=None, num_classes=None, one_hot_encoded=False):
    """attribute.py:111-160."""
    def fn(output):
        if class_idx is not None:
            mask = torch.zeros_like(output); mask[..., class_idx] = 1
        else:
            mask = torch.repeat_interleave(torch.eye(num_classes).to(output), output.size(0) // num_classes, dim=0)
        return mask if one_hot_encoded else output * mask
    return fn


# --------------------------------------------------------------------------- models of SURVEY 8(d)
def toy_model(seed: int = 0, last: int = 64):
    """cfg 1: 5 x [Conv3x3 - ReLU - MaxPool2] on 64x64, widths (8,8,16,16,last), no BN, head
    Linear(4*last,64)-ReLU-Linear(64,64)-ReLU-Linear(64,2): layer indices match LRP_NAME_MAP_TOY."""
    from cxai.model.create_model import VGGType
    torch.manual_seed(seed)
    net = VGGType(n_filters=[8, 8, 16, 16, last], pool_kernels=[(2, 2)] * 5, n_dense=64, n_classes=2, dropout=0.0,
                  block_depth=1, dense_depth=2, input_size=(64, 64), conv_bn=False, dense_bn=False)
    return net.eval()


def genre_model(seed: int = 0, last: int = 256, input_size=(128, 256), randomize_bn: bool = True):
    """cfg 2: arch A widened, filters (64,64,100,128,last), depth 2, BN, pools ((2,4),(2,2)x4), n_dense 100."""
    from cxai.model.create_model import VGGType
    torch.manual_seed(seed)
    net = VGGType(n_filters=[64, 64, 100, 128, last], pool_kernels=[(2, 4), (2, 2), (2, 2), (2, 2), (2, 2)], n_dense=100,
                  n_classes=10, dropout=0.3, block_depth=2, dense_depth=2, input_size=input_size, conv_bn=True,
                  dense_bn=True)
    if randomize_bn:      # non-trivial running statistics so that the BN fold is actually exercised
        g = torch.Generator().manual_seed(seed + 1)
        for m in net.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.data.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
                m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    return net.eval()


def synth_logmel(N: int, H: int, W: int, seed: int):
    """x = clamp(1.2*randn - 1.5, min=-4), [N,1,H,W] (value range of utils/dataloading.py:159-161)."""
    g = torch.Generator().manual_seed(seed)
    return (1.2 * torch.randn(N, 1, H, W, generator=g) - 1.5).clamp(min=-4.0)
