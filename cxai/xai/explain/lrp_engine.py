"""LRP engine: forward pass of the log-mel CNN and the rule-modified backward pass on the CUDA library.

This replaces what zennit does for the reference (``Gradient(model, composite)`` + hooks + autograd,
see preprocessing.py:143-167 and attribute.py:98-107): the layer list of ``model.features`` /
``model.classifier`` is compiled once into a plan of kernel calls, the forward keeps exactly the
activations the backward needs, and the backward stops at the split layer when only the maps there are
wanted (the reference propagates to the input even then).

Semantics: rules return ``input * gradient`` so the quantity flowing between layers
is the relevance itself; layers without a rule (ReLU, MaxPool2d, Dropout, flatten, canonised BatchNorm)
use ordinary autograd: ReLU masks by ``output > 0``, MaxPool routes to the arg-max.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.nn as nn

from drsa_audio_b200 import _lib as _L
from cxai.xai.explain import rules as R
from cxai.model import modify_model as _MM

__all__ = ["LRPPlan", "lrp_intermediate", "lrp_input_relevance", "forward_logits"]


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Op:
    __slots__ = ("kind", "name", "module", "rule", "relu", "w", "b", "w_mod", "wt_mod", "b_mod", "eps", "ones",
                 "kh", "kw", "cin", "cout", "index", "tc")

    def __init__(self, kind, name, module, index):
        self.kind, self.name, self.module, self.index = kind, name, module, index
        self.rule = None
        self.relu = False
        self.tc = None          # lazily prepared tensor-core operands (hi/lo fp16 weights, padded bias)


class LRPPlan:
    """Kernel plan for one (model, composite) pair.  Parameters are prepared once: BatchNorm folded
    (SequentialMergeBatchNorm), rule-modified weights W', b' and the flipped copy for the transposed
    convolution."""

    def __init__(self, model: nn.Module, composite: R.Composite, device: torch.device):
        if model.training:
            raise _L.DRSAError("LRP needs model.eval() (BatchNorm running statistics are folded)")
        self.device = device
        self.use_tc = True          # run model.features on the tcgen05 NHWC pipeline when the shapes allow it
        self.fuse_pool = True       # max-pooling in the epilogue of the convolution before it, where nothing reads the un-pooled map
        self._tc_err = None
        self._tc_ok_cache = {}
        self._graphs = {}                   # replay_pass: key -> "seen" | (graph, static input, static outputs)
        self.ops: List[_Op] = []
        self.module_to_op = {}
        self.filter_index = None          # index of the Projection op of a ProjectionModel
        mods = [(f"features.{n}", m) for n, m in model.features.named_children()]
        mods.append(("flatten", None))
        mods += [(f"classifier.{n}", m) for n, m in model.classifier.named_children()]
        i = 0
        while i < len(mods):
            name, m = mods[i]
            if m is None:
                op = _Op("flatten", name, None, len(self.ops))
            elif isinstance(m, (nn.Conv2d, nn.Linear)):
                op = _Op("conv" if isinstance(m, nn.Conv2d) else "dense", name, m, len(self.ops))
                w = m.weight.detach().to(device, torch.float32)
                b = (m.bias.detach().to(device, torch.float32) if m.bias is not None
                     else torch.zeros(w.shape[0], device=device))
                if isinstance(m, nn.Conv2d):
                    if tuple(m.kernel_size) != (3, 3) or tuple(m.stride) != (1, 1) or m.padding not in ("same", 1, (1, 1)) \
                            or m.groups != 1 or tuple(m.dilation) != (1, 1):
                        raise _L.DRSAError(f"{name}: only 3x3 / stride 1 / 'same' convolutions are on this path")
                nxt = mods[i + 1][1] if i + 1 < len(mods) else None
                if isinstance(nxt, (nn.BatchNorm2d, nn.BatchNorm1d)):
                    if not composite.merges_batchnorm:
                        raise _L.DRSAError(f"{mods[i + 1][0]}: BatchNorm needs the SequentialMergeBatchNorm canonizer")
                    scale = nxt.weight.detach().to(device) / torch.sqrt(nxt.running_var.detach().to(device) + nxt.eps)
                    w = w * scale.view(-1, *([1] * (w.dim() - 1)))
                    b = (b - nxt.running_mean.detach().to(device)) * scale + nxt.bias.detach().to(device)
                    self.module_to_op[nxt] = None          # resolved below: output of BN == output of this op
                op.w, op.b = w.contiguous(), b.contiguous()
                op.cout, op.cin = w.shape[0], w.shape[1]
                op.rule = composite.rule_for(name)
                self._prepare_rule(op)
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                op = _Op("identity", name, m, len(self.ops))       # folded into the previous op
            elif isinstance(m, nn.ReLU):
                op = _Op("relu", name, m, len(self.ops))
            elif isinstance(m, nn.MaxPool2d):
                op = _Op("pool", name, m, len(self.ops))
                ks = m.kernel_size if isinstance(m.kernel_size, (tuple, list)) else (m.kernel_size, m.kernel_size)
                st = m.stride if isinstance(m.stride, (tuple, list)) else (m.stride, m.stride)
                if tuple(st) != tuple(ks) or m.padding not in (0, (0, 0)):
                    raise _L.DRSAError(f"{name}: only non-overlapping max-pooling is on this path")
                op.kh, op.kw = int(ks[0]), int(ks[1])
            elif isinstance(m, nn.Dropout):
                op = _Op("identity", name, m, len(self.ops))
            elif isinstance(m, (_MM.Projection, _MM.SubspaceFilter, _MM.InvProjection)):
                # ProjectionModel (modify_model.py:4-59): h = a U -> SubspaceHook -> a' = h U^T
                kind = {"Projection": "proj", "SubspaceFilter": "sfilter", "InvProjection": "invproj"}[type(m).__name__]
                op = _Op(kind, name, m, len(self.ops))
                op.rule = composite.rule_for(name)
                if kind == "sfilter":
                    if getattr(op.rule, "kind", None) != "subspace_hook":
                        raise _L.DRSAError(f"{name}: expects a SubspaceHook (see explainer.get_class_composite)")
                elif op.rule is None or op.rule.kind != "epsilon":
                    raise _L.DRSAError(f"{name}: expects the Epsilon rule (see explainer.get_class_composite)")
                if kind == "proj":
                    op.w = m.U.detach().to(device, torch.float32).contiguous()          # U [d, m]
                    self.filter_index = len(self.ops)
            else:
                raise _L.DRSAError(f"{name}: layer type {type(m).__name__} is not on this path")
            self.ops.append(op)
            if m is not None:
                self.module_to_op[m] = op
            i += 1
        if self.filter_index is not None:
            f = self.filter_index
            kinds = [o.kind for o in self.ops[f:f + 3]]
            if kinds != ["proj", "sfilter", "invproj"]:
                raise _L.DRSAError("Projection, SubspaceFilter and InvProjection must be consecutive layers")
            self.filter_K = self.ops[f + 1].rule.num_concepts
            if self.ops[f].w.shape[1] % self.filter_K != 0:
                raise _L.DRSAError("num_concepts must divide the number of columns of U")
        # fuse ReLU into the producing conv/dense: the pre-activation is never needed (rules recompute z')
        for k, op in enumerate(self.ops):
            if op.kind in ("conv", "dense"):
                j = k + 1
                while j < len(self.ops) and self.ops[j].kind == "identity":
                    j += 1
                if j < len(self.ops) and self.ops[j].kind == "relu":
                    op.relu = True

    def _prepare_rule(self, op: _Op) -> None:
        rule, w, b = op.rule, op.w, op.b
        op.ones, op.eps = 0, 0.0
        op.w_mod = op.b_mod = op.wt_mod = None
        if rule is None or rule.kind == "pass":
            return
        op.eps = rule.stabilizer
        if rule.kind == "epsilon":
            wm, bm = w, b
        elif rule.kind == "gamma":
            wm, bm = w + rule.gamma * w.clamp(min=0), b + rule.gamma * b.clamp(min=0)
        elif rule.kind == "zplus":
            wm, bm = w.clamp(min=0), b.clamp(min=0)
        elif rule.kind == "wsquare":
            wm, bm, op.ones = w * w, b * b, 1
        elif rule.kind == "flat":
            wm, bm, op.ones = torch.ones_like(w), torch.ones_like(b), 1
        else:
            raise _L.DRSAError(f"{op.name}: rule {rule!r} is not implemented")
        if op.kind == "dense" and rule.kind not in ("epsilon", "gamma", "zplus"):
            raise _L.DRSAError(f"{op.name}: rule {rule!r} is only implemented for convolutions")
        op.w_mod, op.b_mod = wm.contiguous(), bm.contiguous()
        if op.kind == "conv":
            op.wt_mod = torch.empty(op.cin, op.cout, 3, 3, device=w.device)
            _L.check(_L.lib().lrp_conv3x3_flip_weights(_ptr(op.w_mod), op.cout, op.cin, _ptr(op.wt_mod), _stream()),
                     "lrp_conv3x3_flip_weights")

    # ------------------------------------------------------------------ tensor-core conv stack (NHWC)
    @staticmethod
    def _pad64(c: int) -> int:
        return (c + 63) // 64 * 64

    def _stack_end(self) -> int:
        for k, op in enumerate(self.ops):
            if op.kind == "flatten":
                return k
        return len(self.ops)

    def _tc_stack_ok(self, x: torch.Tensor) -> bool:
        """True if every layer of model.features can run on the NHWC tcgen05 pipeline for this input."""
        if not self.use_tc or x.dim() != 4 or x.size(1) != 1:
            return False
        key = tuple(x.shape)
        if key in self._tc_ok_cache:
            return self._tc_ok_cache[key]
        lib = _L.lib()
        B, _, H, W = x.shape
        ok, convs = True, 0
        for k, op in enumerate(self.ops[: self._stack_end()]):
            if op.kind == "conv":
                if k == 0:
                    ok = op.cin == 1
                else:
                    cin_p, cout_p = self._pad64(op.cin), self._pad64(op.cout)
                    ok = cin_p <= 256 and cout_p <= 256 and lib.lrp_tc_conv3x3_supported(B, cin_p, cout_p, H, W) == 0
                    if op.rule is not None and op.rule.kind not in ("epsilon", "gamma", "zplus"):
                        ok = False
                convs += 1
            elif op.kind == "pool":
                ok = H % op.kh == 0 and W % op.kw == 0 and op.kh * op.kw <= 255
                H, W = H // max(op.kh, 1), W // max(op.kw, 1)
            elif op.kind not in ("identity", "relu", "proj", "sfilter", "invproj"):
                ok = False
            if not ok:
                break
        ok = ok and convs >= 2 and self.ops[0].kind == "conv"
        self._tc_ok_cache[key] = ok
        return ok

    def _split_planes(self, t: torch.Tensor):
        hi = torch.empty(t.shape, dtype=torch.float16, device=t.device)
        lo = torch.empty_like(hi)
        _L.check(_L.lib().lrp_tc_split_f16(_ptr(t), t.numel(), _ptr(hi), _ptr(lo), _stream()), "lrp_tc_split_f16")
        return hi, lo

    def _prepare_tc(self, op: _Op, first: bool) -> None:
        """fp16 hi/lo operand planes of a conv layer: forward weights [9][Cout_p][Cin_p]; for the LRP backward the
        rule-modified weights in the same layout and their flipped, channel-swapped copy [9][Cin_p][Cout_p]."""
        if op.tc is not None:
            return
        dev = op.w.device
        cout_p = self._pad64(op.cout)

        def padded_bias(b):
            out = torch.zeros(cout_p, device=dev)
            out[: op.cout] = b
            return out

        if first:
            op.tc = {"w": op.w.reshape(op.cout, 9).contiguous(), "b": padded_bias(op.b), "cout_p": cout_p}
            return
        cin_p = self._pad64(op.cin)

        def taps(w):                     # [Cout, Cin, 3, 3] -> [9, Cout_p, Cin_p], tap = ky*3 + kx
            t = torch.zeros(9, cout_p, cin_p, device=dev)
            t[:, : op.cout, : op.cin] = w.permute(2, 3, 0, 1).reshape(9, op.cout, op.cin)
            return t

        hi, lo = self._split_planes(taps(op.w))
        op.tc = {"hi": hi, "lo": lo, "b": padded_bias(op.b), "cin_p": cin_p, "cout_p": cout_p}
        if op.w_mod is not None:
            mh, ml = self._split_planes(taps(op.w_mod))
            tt = torch.zeros(9, cin_p, cout_p, device=dev)      # transposed conv: tap' = 8 - tap, channels swapped
            tt[:, : op.cin, : op.cout] = op.w_mod.permute(2, 3, 1, 0).reshape(9, op.cin, op.cout).flip(0)
            th, tl = self._split_planes(tt)
            op.tc.update({"m_hi": mh, "m_lo": ml, "m_b": padded_bias(op.b_mod), "t_hi": th, "t_lo": tl})

    def _project_rows(self, op: _Op, a_rows: torch.Tensor, P: int, d: int, ld: int):
        """h = a U and a' = h U^T for position vectors stored as rows [P, ld] (lrp_subspace_project)."""
        U = op.w
        m = U.shape[1]
        h = torch.empty(P, m, device=a_rows.device)
        a_rec = torch.empty(P, ld, device=a_rows.device)
        _L.check(_L.lib().lrp_subspace_project(_ptr(a_rows), _ptr(U), P, d, m, ld, _ptr(h), _ptr(a_rec), _stream()), op.name)
        return h, a_rec

    def _nhwc_to_nchw(self, t):
        """(hi, lo, C, Cp, H, W) NHWC planes -> NCHW fp32 [B, C, H, W]."""
        hi, lo, C, Cp, H, W = t
        B = hi.size(0)
        out = torch.empty(B, C, H, W, device=hi.device)
        _L.check(_L.lib().lrp_tc_nhwc_to_nchw(_ptr(hi), _ptr(lo), B, H, W, Cp, C, _ptr(out), _stream()), "nhwc_to_nchw")
        return out

    def _fusable_pool(self, k: int, keep_from: int, out_index, B: int, H: int, W: int):
        """Index of the MaxPool2d that can be fused into the epilogue of conv op k, or None.  Conditions: only ReLU /
        folded BatchNorm / Dropout between them, nobody reads the un-pooled activation (it is not the requested output,
        and no ReLU in between needs its mask in the backward), and the kernel covers the shape."""
        if out_index is None or k == 0 or not self.fuse_pool:
            return None
        op = self.ops[k]
        j = k + 1
        end = self._stack_end()
        while j < end and self.ops[j].kind in ("identity", "relu"):
            if self.ops[j].kind == "relu" and j >= keep_from and self._relu_needs_mask(j):
                return None
            j += 1
        if j >= end or self.ops[j].kind != "pool" or k <= out_index < j:
            return None
        pool = self.ops[j]
        if _L.lib().lrp_tc_conv3x3_pool_supported(B, self._pad64(op.cin), self._pad64(op.cout), H, W, pool.kh, pool.kw) != 0:
            return None
        return j

    def _forward_tc_stack(self, x: torch.Tensor, keep_from: int, saved, outs, out_index=None):
        """model.features on the tensor cores.  Keeps, for ops >= keep_from, the NHWC input planes of every conv and
        the arg-max of every pool; returns the features as NCHW fp32 for the dense head.  ``out_index``: the only op
        whose output the caller reads from ``outs`` (-1: none; None: any, which disables the conv + pool fusion)."""
        lib = _L.lib()
        B, _, H, W = x.shape
        dev = x.device
        if self._tc_err is None or self._tc_err.device != dev:
            self._tc_err = torch.zeros(1, dtype=torch.int32, device=dev)
        end = self._stack_end()
        cur = None                                   # (hi, lo, C, Cp, H, W)
        skip_until = -1
        for k in range(end):
            op = self.ops[k]
            keep = k >= keep_from
            if k <= skip_until:                      # ReLU / pool already done by the fused conv + pool kernel
                if op.kind == "relu" and keep:
                    saved[k] = ("tc_relu", None)     # never dereferenced: _fusable_pool checked that no mask is needed
                outs[k] = ("nhwc", cur) if (k == skip_until and k >= keep_from - 1) else None
                continue
            kp = self._fusable_pool(k, keep_from, out_index, B, H, W) if op.kind == "conv" else None
            if kp is not None:
                pool = self.ops[kp]
                self._prepare_tc(op, first=False)
                cout_p = op.tc["cout_p"]
                Ho, Wo = H // pool.kh, W // pool.kw
                yh = torch.empty(B, Ho, Wo, cout_p, dtype=torch.float16, device=dev)
                yl = torch.empty_like(yh)
                am = torch.empty(B, Ho, Wo, cout_p, dtype=torch.uint8, device=dev) if kp >= keep_from else None
                _L.check(lib.lrp_tc_conv3x3_forward_pool(_ptr(cur[0]), _ptr(cur[1]), _ptr(op.tc["hi"]), _ptr(op.tc["lo"]),
                                                         _ptr(op.tc["b"]), B, H, W, op.tc["cin_p"], cout_p, op.cout,
                                                         int(op.relu), pool.kh, pool.kw, _ptr(yh), _ptr(yl), _ptr(am),
                                                         _ptr(self._tc_err), _stream()), op.name)
                if keep:
                    saved[k] = ("tc_conv", cur)
                if kp >= keep_from:
                    saved[kp] = ("tc_pool", am, (H, W, cout_p))
                H, W = Ho, Wo
                cur = (yh, yl, op.cout, cout_p, H, W)
                skip_until = kp
                outs[k] = None
                continue
            if op.kind == "conv":
                self._prepare_tc(op, first=(k == 0))
                cout_p = op.tc["cout_p"]
                yh = torch.empty(B, H, W, cout_p, dtype=torch.float16, device=dev)
                yl = torch.empty_like(yh)
                if k == 0:
                    _L.check(lib.lrp_tc_conv3x3_first(_ptr(x), _ptr(op.tc["w"]), _ptr(op.tc["b"]), B, H, W, op.cout, cout_p,
                                                      int(op.relu), _ptr(yh), _ptr(yl), _stream()), op.name)
                    if keep:
                        saved[k] = ("tc_first", x)
                else:
                    _L.check(lib.lrp_tc_conv3x3_forward(_ptr(cur[0]), _ptr(cur[1]), _ptr(op.tc["hi"]), _ptr(op.tc["lo"]),
                                                        _ptr(op.tc["b"]), B, H, W, op.tc["cin_p"], cout_p, op.cout,
                                                        int(op.relu), _ptr(yh), _ptr(yl), None, _ptr(self._tc_err),
                                                        _stream()), op.name)
                    if keep:
                        saved[k] = ("tc_conv", cur)
                cur = (yh, yl, op.cout, cout_p, H, W)
            elif op.kind == "pool":
                Ho, Wo = H // op.kh, W // op.kw
                C, Cp = cur[2], cur[3]
                yh = torch.empty(B, Ho, Wo, Cp, dtype=torch.float16, device=dev)
                yl = torch.empty_like(yh)
                am = torch.empty(B, Ho, Wo, Cp, dtype=torch.uint8, device=dev) if keep else None
                _L.check(lib.lrp_tc_maxpool(_ptr(cur[0]), _ptr(cur[1]), B, H, W, Cp, op.kh, op.kw, _ptr(yh), _ptr(yl),
                                            _ptr(am), _stream()), op.name)
                if keep:
                    saved[k] = ("tc_pool", am, (H, W, Cp))
                H, W = Ho, Wo
                cur = (yh, yl, C, Cp, H, W)
            elif op.kind == "relu":
                if keep:
                    saved[k] = ("tc_relu", cur)          # ReLU is fused into the producer: output == input
            elif op.kind == "proj":
                C, Cp = cur[2], cur[3]
                P = B * H * W
                a_rows = torch.empty(P, Cp, device=dev)
                _L.check(lib.lrp_tc_planes_to_f32(_ptr(cur[0]), _ptr(cur[1]), a_rows.numel(), _ptr(a_rows), _stream()), op.name)
                h, a_rec = self._project_rows(op, a_rows, P, C, Cp)
                saved[k] = ("tc_filter", a_rows, h, a_rec, (B, H, W, C, Cp))
                yh = torch.empty(B, H, W, Cp, dtype=torch.float16, device=dev)
                yl = torch.empty_like(yh)
                _L.check(lib.lrp_tc_split_f16(_ptr(a_rec), a_rec.numel(), _ptr(yh), _ptr(yl), _stream()), op.name)
                cur = (yh, yl, C, Cp, H, W)
            outs[k] = ("nhwc", cur) if k >= keep_from - 1 else None
        return self._nhwc_to_nchw(cur)

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, keep_from: int = 0, out_index=None):
        """Runs the network; returns (logits, saved, outs) where saved[k] holds what op k needs in the backward
        (its input, and the arg-max for pooling).  Only ops with index >= keep_from are kept.  ``out_index``: see
        _forward_tc_stack."""
        lib = _L.lib()
        saved = [None] * len(self.ops)
        outs = [None] * len(self.ops)
        cur = x
        n_tc = 0
        if self._tc_stack_ok(x):
            n_tc = self._stack_end()
            cur = self._forward_tc_stack(x, keep_from, saved, outs, out_index)
        for k, op in enumerate(self.ops):
            if k < n_tc:
                continue
            keep = k >= keep_from
            if op.kind == "conv":
                N, Cin, H, W = cur.shape
                y = torch.empty(N, op.cout, H, W, device=cur.device)
                _L.check(lib.lrp_conv3x3_forward(_ptr(cur), _ptr(op.w), _ptr(op.b), N, Cin, op.cout, H, W,
                                                 int(op.relu), _ptr(y), _stream()), op.name)
                if keep:
                    saved[k] = cur
                cur = y
            elif op.kind == "dense":
                N, In = cur.shape
                y = torch.empty(N, op.cout, device=cur.device)
                _L.check(lib.lrp_dense_forward(_ptr(cur), _ptr(op.w), _ptr(op.b), N, In, op.cout, int(op.relu),
                                               _ptr(y), _stream()), op.name)
                if keep:
                    saved[k] = cur
                cur = y
            elif op.kind == "pool":
                N, Cc, H, W = cur.shape
                Ho, Wo = H // op.kh, W // op.kw
                y = torch.empty(N, Cc, Ho, Wo, device=cur.device)
                am = torch.empty(N, Cc, Ho, Wo, dtype=torch.int32, device=cur.device)
                _L.check(lib.lrp_maxpool_forward(_ptr(cur), N * Cc, H, W, op.kh, op.kw, _ptr(y), _ptr(am),
                                                 _stream()), op.name)
                if keep:
                    saved[k] = (am, (N, Cc, H, W))
                cur = y
            elif op.kind == "flatten":
                if keep:
                    saved[k] = ("flat", cur.shape, cur)      # the features also bound |s| of the conv stack below
                cur = cur.reshape(cur.size(0), -1)
            elif op.kind == "relu":
                if keep:
                    saved[k] = cur            # ReLU already applied by the producer: output == input here
            elif op.kind == "proj":
                N, Cc, H, W = cur.shape
                P = N * H * W
                a_rows = torch.empty(P, Cc, device=cur.device)
                _L.check(lib.lrp_tc_nchw_to_nhwc_f32(_ptr(cur.contiguous()), N, H, W, Cc, Cc, _ptr(a_rows), _stream()), op.name)
                h, a_rec = self._project_rows(op, a_rows, P, Cc, Cc)
                saved[k] = ("filter", a_rows, h, a_rec, (N, H, W, Cc, Cc))
                y = torch.empty(N, Cc, H, W, device=cur.device)
                _L.check(lib.lrp_tc_nhwc_f32_to_nchw(_ptr(a_rec), N, H, W, Cc, Cc, _ptr(y), _stream()), op.name)
                cur = y
            outs[k] = cur
        return cur, saved, outs

    # ------------------------------------------------------------------ backward
    def backward(self, Rel: torch.Tensor, saved, stop_after: int = -1, start: Optional[int] = None, nhwc=None,
                 bound_feat: Optional[torch.Tensor] = None, keep_nhwc: bool = False) -> torch.Tensor:
        """Propagates relevance from the logits (or, with ``start``, from the output of op ``start``) down to the OUTPUT
        of op ``stop_after`` (-1: to the input).  Inside the tensor-core conv stack the relevance travels as NHWC fp32
        with padded channels (``nhwc`` = (C, Cp, H, W) when ``Rel`` already is).  ``bound_feat``: activation with the
        layout of ``Rel`` such that Rel = bound_feat * c; max |c| per sample bounds the quotients of the layer below.
        A ProjectionModel's filter turns the batch of B samples into B*(K+1) relevance maps (clone order of the
        reference: sample-major)."""
        lib = _L.lib()
        feat = None            # output of model.features (NCHW fp32), set when the flatten op is crossed
        bound = None           # per-sample bound on |s| of the next tensor-core layer (device, [B])
        cmax = None
        if nhwc is not None:
            cmax = torch.zeros(len(self.ops) + 1, Rel.size(0), device=Rel.device)
            if bound_feat is not None:
                bound = cmax[len(self.ops)]
                _L.check(lib.lrp_tc_sample_absmax_ratio(_ptr(Rel), _ptr(bound_feat), Rel.size(0), Rel[0].numel(), _ptr(bound),
                                                        _stream()), "absmax_ratio")
        for k in range(len(self.ops) - 1 if start is None else start, stop_after, -1):
            op = self.ops[k]
            sv = saved[k]
            if op.kind in ("sfilter", "proj"):
                continue
            if op.kind == "invproj":
                return self._filter_backward(Rel, saved, stop_after, nhwc)
            is_tc = isinstance(sv, tuple) and len(sv) > 0 and isinstance(sv[0], str) and sv[0].startswith("tc_")
            if is_tc and nhwc is None and sv[0] != "tc_first":
                # entering the NHWC stack from the dense head: Rel is NCHW fp32 [B, C, H, W]
                B, C, H, W = Rel.shape
                Cp = self._pad64(C)
                Rel = Rel.contiguous()
                cmax = torch.zeros(len(self.ops) + 1, B, device=Rel.device)
                if feat is not None and feat.shape == Rel.shape:
                    # R = a * c at the output of the stack: max |c| bounds |s| of the first rule layer below
                    bound = cmax[len(self.ops)]
                    _L.check(lib.lrp_tc_sample_absmax_ratio(_ptr(Rel), _ptr(feat), B, C * H * W, _ptr(bound), _stream()),
                             "absmax_ratio")
                t = torch.empty(B, H, W, Cp, device=Rel.device)
                _L.check(lib.lrp_tc_nchw_to_nhwc_f32(_ptr(Rel), B, H, W, C, Cp, _ptr(t), _stream()), "to_nhwc")
                Rel, nhwc = t, (C, Cp, H, W)
            if is_tc and sv[0] == "tc_pool":
                am, (H, W, Cp) = sv[1], sv[2]
                B = Rel.size(0)
                R_in = torch.empty(B, H, W, Cp, device=Rel.device)
                _L.check(lib.lrp_tc_maxpool_backward(_ptr(Rel), _ptr(am), B, H, W, Cp, op.kh, op.kw, _ptr(R_in), _stream()),
                         op.name)
                Rel, nhwc = R_in, (nhwc[0], Cp, H, W)
                continue
            if is_tc and sv[0] == "tc_relu":
                if self._relu_needs_mask(k):
                    a = sv[1]
                    _L.check(lib.lrp_tc_relu_mask(_ptr(Rel), _ptr(a[0]), _ptr(a[1]), Rel.numel(), _stream()), op.name)
                continue
            if is_tc and sv[0] == "tc_conv":
                if op.rule is None:
                    raise _L.DRSAError(f"{op.name}: no LRP rule assigned")
                xin = sv[1]                               # (hi, lo, C, Cp, H, W) input planes
                B, H, W = Rel.size(0), xin[4], xin[5]
                cin_p, cout_p = op.tc["cin_p"], op.tc["cout_p"]
                sh = torch.empty(B, H, W, cout_p, dtype=torch.float16, device=Rel.device)
                sl = torch.empty_like(sh)
                _L.check(lib.lrp_tc_conv3x3_ratio(_ptr(xin[0]), _ptr(xin[1]), _ptr(op.tc["m_hi"]), _ptr(op.tc["m_lo"]),
                                                  _ptr(op.tc["m_b"]), _ptr(Rel), B, H, W, cin_p, cout_p, op.eps, _ptr(bound),
                                                  _ptr(sh), _ptr(sl), _ptr(self._tc_err), _stream()), op.name)
                R_in = torch.empty(B, H, W, cin_p, device=Rel.device)
                nxt = cmax[k]
                _L.check(lib.lrp_tc_conv3x3_inputmul(_ptr(sh), _ptr(sl), _ptr(op.tc["t_hi"]), _ptr(op.tc["t_lo"]), _ptr(xin[0]),
                                                     _ptr(xin[1]), B, H, W, cout_p, cin_p, _ptr(bound), _ptr(nxt), _ptr(R_in),
                                                     _ptr(self._tc_err), _stream()), op.name)
                Rel, nhwc, bound = R_in, (op.cin, cin_p, H, W), nxt
                continue
            if is_tc and sv[0] == "tc_first" and nhwc is not None and op.ones and nhwc[1] == 64 and op.rule is not None \
                    and op.rule.kind in ("wsquare", "flat"):
                # WSquare / Flat on the first layer: one pass over the NHWC relevance (no layout conversion, no s buffer)
                B, H, W = Rel.size(0), nhwc[2], nhwc[3]
                R_in = torch.empty(B, 1, H, W, device=Rel.device)
                _L.check(lib.lrp_tc_first_ones_backward(_ptr(Rel), _ptr(op.w_mod), _ptr(op.b_mod), B, H, W, op.cout, nhwc[1],
                                                        op.eps, _ptr(R_in), _stream()), op.name)
                Rel, nhwc = R_in, None
                continue
            if is_tc and sv[0] == "tc_first":
                # leave the NHWC stack: the first conv (Cin = 1) runs on the CUDA-core kernels in NCHW
                if nhwc is not None:
                    Rel = self._rel_to_nchw(Rel, nhwc)
                    nhwc = None
                sv = sv[1]
            if op.kind == "dense":
                if op.rule is None:
                    raise _L.DRSAError(f"{op.name}: no LRP rule assigned")
                if op.rule.kind == "pass":
                    continue
                x = sv
                N, In = x.shape
                s_buf = torch.empty(N, op.cout, device=x.device)
                R_in = torch.empty_like(x)
                _L.check(lib.lrp_dense_epsilon_backward(_ptr(x), _ptr(op.w_mod), _ptr(op.b_mod), _ptr(Rel), N, In,
                                                        op.cout, op.eps, _ptr(s_buf), _ptr(R_in), _stream()), op.name)
                Rel = R_in
            elif op.kind == "conv":
                if op.rule is None:
                    raise _L.DRSAError(f"{op.name}: no LRP rule assigned")
                if op.rule.kind == "pass":
                    continue
                x = sv
                N, Cin, H, W = x.shape
                s_buf = torch.empty(N, op.cout, H, W, device=x.device)
                R_in = torch.empty_like(x)
                _L.check(lib.lrp_conv3x3_backward(_ptr(x), _ptr(op.w_mod), _ptr(op.wt_mod), _ptr(op.b_mod), _ptr(Rel), N,
                                                  Cin, op.cout, H, W, op.eps, op.ones, _ptr(s_buf), _ptr(R_in),
                                                  _stream()), op.name)
                Rel = R_in
            elif op.kind == "relu":
                a = sv
                Rel = Rel.contiguous()
                _L.check(lib.lrp_relu_mask(_ptr(a), _ptr(Rel), Rel.numel(), _stream()), op.name)
            elif op.kind == "pool":
                am, shp = sv
                N, Cc, H, W = shp
                R_in = torch.empty(shp, device=Rel.device)
                _L.check(lib.lrp_maxpool_backward(_ptr(Rel.contiguous()), _ptr(am), N * Cc, H, W, op.kh, op.kw,
                                                  _ptr(R_in), _stream()), op.name)
                Rel = R_in
            elif op.kind == "flatten":
                Rel = Rel.reshape(sv[1])
                feat = sv[2]
        if keep_nhwc:            # (relevance, (C, Cp, H, W) or None): the caller reads the NHWC fp32 relevance itself
            return Rel, nhwc
        if nhwc is not None:
            Rel = self._rel_to_nchw(Rel, nhwc)
        return Rel

    def _filter_backward(self, Rel: torch.Tensor, saved, stop_after: int, nhwc) -> torch.Tensor:
        """InvProjection (Epsilon) -> SubspaceHook -> Projection (Epsilon) for all K+1 clones at once, then the layers
        below once per clone on the shared forward state.  Returns [B*(K+1), ...] in the reference's clone order."""
        lib = _L.lib()
        f = self.filter_index
        if stop_after >= f - 1:
            raise _L.DRSAError("relevance can only be read out below the projection layers of a ProjectionModel")
        mode, a_rows, h, a_rec, (B, H, W, C, ld) = saved[f]
        K, U = self.filter_K, self.ops[f].w
        m = U.shape[1]
        P = B * H * W
        if mode == "tc_filter":
            if nhwc is None:                       # relevance arrives NCHW (no conv layer above the split)
                t = torch.empty(B, H, W, ld, device=Rel.device)
                _L.check(lib.lrp_tc_nchw_to_nhwc_f32(_ptr(Rel.contiguous()), B, H, W, C, ld, _ptr(t), _stream()), "to_nhwc")
                Rel = t
        else:
            if nhwc is not None:
                Rel = self._rel_to_nchw(Rel, nhwc)
            t = torch.empty(B, H, W, ld, device=Rel.device)
            _L.check(lib.lrp_tc_nchw_to_nhwc_f32(_ptr(Rel.contiguous()), B, H, W, C, ld, _ptr(t), _stream()), "to_nhwc")
            Rel = t
        out = torch.empty(K + 1, B, H, W, ld, device=Rel.device)
        ws = torch.empty(int(_L.check(lib.lrp_subspace_filter_workspace_bytes(P, C, m))), dtype=torch.uint8, device=Rel.device)
        _L.check(lib.lrp_subspace_filter(_ptr(a_rows), _ptr(h), _ptr(a_rec), _ptr(Rel.contiguous()), _ptr(U), P, C, m, K, ld,
                                         self.ops[f + 2].rule.stabilizer, self.ops[f].rule.stabilizer, _ptr(out), _ptr(ws),
                                         ws.numel(), _stream()), "lrp_subspace_filter")
        res = []
        for kk in range(K + 1):
            if mode == "tc_filter":
                res.append(self.backward(out[kk], saved, stop_after, start=f - 1, nhwc=(C, ld, H, W),
                                         bound_feat=a_rows))
            else:
                Rk = torch.empty(B, C, H, W, device=Rel.device)
                _L.check(lib.lrp_tc_nhwc_f32_to_nchw(_ptr(out[kk]), B, H, W, ld, C, _ptr(Rk), _stream()), "to_nchw")
                res.append(self.backward(Rk, saved, stop_after, start=f - 1))
        return torch.stack(res, dim=1).flatten(0, 1)

    def _rel_to_nchw(self, Rel: torch.Tensor, nhwc) -> torch.Tensor:
        C, Cp, H, W = nhwc
        B = Rel.size(0)
        out = torch.empty(B, C, H, W, device=Rel.device)
        _L.check(_L.lib().lrp_tc_nhwc_f32_to_nchw(_ptr(Rel), B, H, W, Cp, C, _ptr(out), _stream()), "to_nchw")
        return out

    def _relu_needs_mask(self, k: int) -> bool:
        """The mask of an un-hooked ReLU is a no-op when the relevance arriving at its output was produced by an
        input-multiplying rule on that very output (Gamma/ZPlus/Epsilon return x * (...), and pooling only routes
        such values): then R is already zero wherever the activation is zero."""
        for j in range(k + 1, len(self.ops)):
            o = self.ops[j]
            if o.kind in ("identity", "pool", "flatten", "relu"):
                continue
            if o.kind == "proj":
                return False          # Epsilon on the projection: R = a * (...)
            if o.kind in ("conv", "dense"):
                return not (o.rule is not None and o.rule.kind in ("epsilon", "gamma", "zplus"))
            return True
        return True

    def check_nonneg_inputs(self):
        """Gamma / ZPlus use the collapsed form that is valid for non-negative inputs only: every such
        layer must be fed by a ReLU / pooling output (true for every name map of the reference)."""
        seen_relu = False
        for op in self.ops:
            if op.kind in ("conv", "dense") and op.rule is not None and op.rule.kind in ("gamma", "zplus"):
                if not seen_relu:
                    raise _L.DRSAError(f"{op.name}: Gamma/ZPlus on a layer with signed input is not on this path "
                                       "(use WSquare/Flat/Epsilon for the first layer, as the reference does)")
                # zennit switches to the W + gamma*W^- branch where the layer's OUTPUT is negative; the collapsed form
                # assumes the relevance arriving there is zero, which only a ReLU right behind the layer guarantees
                if not op.relu:
                    raise _L.DRSAError(f"{op.name}: Gamma/ZPlus needs a ReLU directly behind the layer (BatchNorm / Dropout "
                                       "in between are fine); the negative-output branch of the rule is not on this path")
            if op.kind == "relu" or (op.kind in ("conv", "dense") and op.relu):
                seen_relu = True
            elif op.kind in ("conv", "dense"):
                seen_relu = False


    # ------------------------------------------------------------------ CUDA-graph replay of whole engine passes
    def replay_pass(self, key, inputs, body):
        """``body(*inputs)`` -> tuple of tensors, as ONE CUDA-graph launch from the third call with the same key on
        (``inputs``: a tensor or a tuple of tensors; everything that changes between calls must be in it).

        An engine pass is ~180 launches (35 layers forward and backward, plus the fill / copy kernels of the tensors it
        allocates); launched one by one they leave ~0.25 ms of gaps per 256 samples and keep the host busy.  A pass
        has no host synchronisation and no data-dependent control flow, so it is captured once per (pass kind, input
        shape, split layer, seed, arithmetic) into a graph whose private memory pool keeps every intermediate at a fixed
        address; a replay copies the batch into the static input and the results out of the static outputs.  First
        call with a key: plain launches (one-time initialisations must not happen inside a capture); second call:
        capture; at most ``GRAPH_CACHE`` graphs are kept (each pins the intermediates of one pass, ~46 MB per sample of
        128 x 256)."""
        if isinstance(inputs, torch.Tensor):
            inputs = (inputs,)
        entry = self._graphs.get(key)
        if entry is None or entry == "never":
            out = body(*inputs)
            self._graphs[key] = "seen" if out is not None and entry is None else "never"
            return out
        if entry == "seen":
            while len([v for v in self._graphs.values() if isinstance(v, tuple)]) >= GRAPH_CACHE:
                oldest = next(k for k, v in self._graphs.items() if isinstance(v, tuple))
                del self._graphs[oldest]
            static_in = tuple(t.clone() for t in inputs)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = body(*static_in)
            entry = (graph, static_in, outs)
            self._graphs.pop(key, None)
            self._graphs[key] = entry                  # (re)inserted last: the dict order is the age
        else:
            for dst, src in zip(entry[1], inputs):
                dst.copy_(src)
        entry[0].replay()
        return tuple(o.clone() for o in entry[2])      # the static outputs are overwritten by the next replay

    def tc_failed(self) -> bool:
        """True (and the tensor-core stack is switched off for this plan) if a tensor-core kernel raised its error
        flag: a value left the fp16 range of the split planes.  The caller reruns on the fp32 CUDA-core kernels."""
        if self._tc_err is None or not self.use_tc:
            return False
        code = int(self._tc_err.item())
        if code == 0:
            return False
        if code == 1:
            raise _L.DRSAError("tensor-core convolution: misaligned shared-memory window")
        import warnings
        warnings.warn("LRP: activations or relevance quotients exceed the fp16 range of the split-precision tensor-core "
                      "path; rerunning this model on the fp32 CUDA-core kernels")
        self.use_tc = False
        self._tc_err.zero_()
        return True



# Samples per engine pass.  The reference cuts the batch into minibatches of ``attr_batch_size`` = 64 to bound autograd
# memory (preprocessing.py:150-167); every kernel here treats samples independently (per-sample scales, eval-mode
# BatchNorm folded), so the result does not depend on the cut and the engine is free to take more samples per pass:
# the last conv layers and the dense head only fill the 148 SMs from ~256 samples on.  ``attr_batch_size`` is honoured
# as a lower bound; the upper bound keeps the activation planes of one pass below ~12 GB.
ENGINE_CHUNK = 256
USE_GRAPH = True             # replay whole engine passes as CUDA graphs (LRPPlan.replay_pass)
GRAPH_MIN_SAMPLES = 32       # small passes are latency-bound either way and not worth pinning memory for
GRAPH_CACHE = 3


def _engine_chunk(x: torch.Tensor, requested: int) -> int:
    per_sample = max(1, x[0].numel()) * 1400          # bytes: hi/lo planes + fp32 relevance of the widest layers
    return max(int(requested), min(ENGINE_CHUNK, max(1, (12 << 30) // per_sample)))


def _fingerprint(model) -> tuple:
    """Cheap identity of everything a plan bakes in: storage and in-place version of every parameter, buffer and
    projection matrix, and the train/eval state.  ``load_state_dict``, an optimiser step or ``model.train()`` change it."""
    parts = [bool(model.training)]
    for t in list(model.parameters()) + list(model.buffers()):
        parts.append((t.data_ptr(), t._version))
    for mod in model.modules():
        for name in ("U", "U_inv"):
            t = mod.__dict__.get(name)
            if isinstance(t, torch.Tensor):
                parts.append((t.data_ptr(), t._version, tuple(t.shape)))
    return tuple(parts)


def _plan(model, composite, device) -> LRPPlan:
    """The compiled plan of (model, composite) on ``device``.  It lives ON the model object (and dies with it: ids of freed
    models / composites can be reused by CPython), holds a strong reference to its composite, and is rebuilt whenever the
    model's fingerprint changes."""
    cache = model.__dict__.setdefault("_drsa_b200_plans", {})
    key = str(device)
    entry = cache.get(key)
    fp = _fingerprint(model)
    if entry is None or entry[1] is not composite or entry[2] != fp:
        p = LRPPlan(model, composite, device)
        p.check_nonneg_inputs()
        cache[key] = entry = (p, composite, fp)
    return entry[0]


def invalidate_plans(model) -> None:
    """Drop the compiled plans of ``model`` (only needed after changes the fingerprint cannot see)."""
    model.__dict__.pop("_drsa_b200_plans", None)


def _prep_input(input_batch: torch.Tensor) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise _L.DRSAError("no CUDA device available and no fallback path exists")
    x = input_batch.detach()
    if not x.is_cuda:
        x = x.cuda(non_blocking=True)
    return x.to(torch.float32).contiguous()


def forward_logits(model, input_batch, composite=None, batch_size: int = 64) -> torch.Tensor:
    x = _prep_input(input_batch)
    batch_size = _engine_chunk(x, batch_size)
    with torch.cuda.device(x.device):
        plan = _plan(model, composite or R.NameMapComposite([], canonizers=[R.SequentialMergeBatchNorm()]), x.device)
        while True:
            outs = [plan.forward(x[i:i + batch_size], keep_from=len(plan.ops), out_index=-1)[0]
                    for i in range(0, x.size(0), batch_size)]
            if not plan.tc_failed():
                break
    return torch.cat(outs, 0)


def lrp_intermediate(model, input_batch, composite, layer, class_idx, attr_batch_size: int = 64,
                     one_hot_encoded: bool = False, attr_output_fn: Optional[Callable] = None):
    """(activation_maps, relevance_maps) at the output of ``layer`` (get_intermediate, preprocessing.py:106-176).
    Minibatches of ``attr_batch_size`` like the reference; the backward stops at ``layer``."""
    from cxai.xai.explain.attribute import lrp_output_modifier
    x = _prep_input(input_batch)
    fn = attr_output_fn or lrp_output_modifier(class_idx, one_hot_encoded=one_hot_encoded)
    if _seed_is_rowwise(fn):          # a seed that looks at the minibatch as a whole keeps the reference's cut
        attr_batch_size = _engine_chunk(x, attr_batch_size)
    with torch.cuda.device(x.device):
        plan = _plan(model, composite, x.device)
        op = plan.module_to_op.get(layer)
        if op is None:
            # a BatchNorm folded into its conv: its output is the conv op's output
            for k, o in enumerate(plan.ops):
                if o.module is layer:
                    op = o
            if op is None:
                raise _L.DRSAError("layer is not a module of model.features / model.classifier")
        split = op.index
        # a conv/dense whose ReLU was fused cannot expose its pre-activation
        if plan.ops[split].kind in ("conv", "dense", "identity") and any(
                o.kind in ("conv", "dense") and o.relu for o in plan.ops[max(0, split - 1):split + 1]):
            nxt = split
            while plan.ops[nxt].kind != "relu":
                nxt += 1
            raise _L.DRSAError(f"split at {plan.ops[split].name} (pre-activation) is not on this path; "
                               f"use the ReLU {plan.ops[nxt].name} like the reference (layers 19/26/33)")
        def one_pass(xb, mask=None):
            logits, saved, outs = plan.forward(xb, keep_from=split + 1, out_index=split)
            seed = fn(logits).contiguous() if mask is None else _class_seed(logits, mask, fn_key[2])
            r = plan.backward(seed, saved, stop_after=split)
            o = outs[split]
            return (plan._nhwc_to_nchw(o[1]) if isinstance(o, tuple) else o), r
        fn_key = getattr(fn, "key", None)
        mask_row = None
        while True:
            a_maps, r_maps = [], []
            for i in range(0, x.size(0), attr_batch_size):
                xb = x[i:i + attr_batch_size]
                if USE_GRAPH and fn_key is not None and xb.size(0) >= GRAPH_MIN_SAMPLES:
                    if mask_row is None:
                        mask_row = torch.zeros(plan.ops[-1].cout, device=x.device)
                        mask_row[fn_key[1]] = 1.0
                    a, r = plan.replay_pass(("intermediate", tuple(xb.shape), split, fn_key[2], plan.use_tc),
                                            (xb, mask_row), one_pass)
                else:
                    a, r = one_pass(xb)
                a_maps.append(a)
                r_maps.append(r)
            if not plan.tc_failed():
                break
    return torch.cat(a_maps, 0), torch.cat(r_maps, 0)


def _class_seed(logits: torch.Tensor, mask_row: torch.Tensor, one_hot: bool) -> torch.Tensor:
    """The seed of ``lrp_output_modifier(class_idx, one_hot_encoded=...)`` (attribute.py:134-144) with the class mask as a
    TENSOR [n_classes] instead of a Python index, so that one captured pass serves every class (``mask_row`` is a static
    input of the graph)."""
    mask = mask_row.to(logits.dtype).expand_as(logits)
    return (mask if one_hot else logits * mask).contiguous()


def lrp_context_pairs(model, input_batch, composite, layer, class_idx, idcs=None, attr_batch_size: int = 64,
                      one_hot_encoded: bool = False):
    """Rows for DRSA straight from the engine: (activation vectors, context vectors c = R / (a + 1e-7)) [N*L, d] at the
    output of ``layer`` and their two sums of squares (float64 [2], un-normalised) -- what ``get_intermediate`` +
    ``get_vectors_from_maps`` + ``compute_context_vectors`` (preprocessing.py:106-256) produce, without materialising the
    NCHW maps: inside the tensor-core stack activations and relevance are NHWC, where a row of the result IS a position.
    ``idcs`` [N, L] selects positions per sample (None: all).  Returns None if the split layer is not inside the NHWC
    stack (the caller then takes the map route)."""
    from cxai.xai.explain.attribute import lrp_output_modifier
    x = _prep_input(input_batch)
    fn = lrp_output_modifier(class_idx, one_hot_encoded=one_hot_encoded)
    bs = _engine_chunk(x, attr_batch_size)
    lib = _L.lib()
    with torch.cuda.device(x.device):
        plan = _plan(model, composite, x.device)
        op = plan.module_to_op.get(layer)
        if op is None or not plan._tc_stack_ok(x[:1]) or op.index >= plan._stack_end() or plan.ops[op.index].kind != "relu":
            return None
        split = op.index
        idx_all = None if idcs is None else torch.as_tensor(idcs, dtype=torch.int64).to(x.device).contiguous()

        mask_row = None

        def one_pass(xb, idx=None, mask=None):
            logits, saved, outs = plan.forward(xb, keep_from=split + 1, out_index=split)
            seed = fn(logits).contiguous() if mask is None else _class_seed(logits, mask, one_hot_encoded)
            Rel, nhwc = plan.backward(seed, saved, stop_after=split, keep_nhwc=True)
            o = outs[split]
            if nhwc is None or not isinstance(o, tuple):
                return None
            hi, lo, C, Cp, H, W = o[1]
            B, HW = xb.size(0), H * W
            L = HW if idx is None else idx.size(1)
            act = torch.empty(B * L, C, device=xb.device)
            ctx = torch.empty_like(act)
            ss = torch.zeros(2, dtype=torch.float64, device=xb.device)
            _L.check(lib.drsa_context_pairs_nhwc(_ptr(hi), _ptr(lo), _ptr(Rel), B, HW, Cp, C, _ptr(idx), L, _ptr(act), _ptr(ctx),
                                                 _ptr(ss), _stream()), "drsa_context_pairs_nhwc")
            return act, ctx, ss
        while True:
            acts, ctxs, total = [], [], torch.zeros(2, dtype=torch.float64, device=x.device)
            ok = True
            for i in range(0, x.size(0), bs):
                xb = x[i:i + bs]
                idx = None if idx_all is None else idx_all[i:i + bs].contiguous()
                if USE_GRAPH and idx is None and xb.size(0) >= GRAPH_MIN_SAMPLES and plan.use_tc:
                    if mask_row is None:
                        n_out = plan.ops[-1].cout
                        mask_row = torch.zeros(n_out, device=x.device)
                        mask_row[class_idx] = 1.0
                    res = plan.replay_pass(("pairs", tuple(xb.shape), split, bool(one_hot_encoded), plan.use_tc),
                                           (xb, mask_row), lambda a, b: one_pass(a, None, b))
                else:
                    res = one_pass(xb, idx)
                if res is None:
                    ok = False
                    break
                acts.append(res[0]); ctxs.append(res[1]); total += res[2]
            if not ok:
                return None
            if not plan.tc_failed():
                break
            if not plan.use_tc:
                return None
    return torch.cat(acts, 0), torch.cat(ctxs, 0), total


def _seed_is_rowwise(fn) -> bool:
    return bool(getattr(fn, "rowwise", False))


def _run_relevance(plan, x, fn, batch_size, seed_rows=None):
    out = []
    fn_key = getattr(fn, "key", None) if seed_rows is None else None
    mask_row = None

    def one_pass(xb, mask):
        logits, saved, _ = plan.forward(xb, keep_from=0, out_index=-1)
        return (plan.backward(_class_seed(logits, mask, fn_key[2]), saved, stop_after=-1),)
    for i in range(0, x.size(0), batch_size):
        xb = x[i:i + batch_size]
        if USE_GRAPH and fn_key is not None and xb.size(0) >= GRAPH_MIN_SAMPLES and plan.filter_index is None:
            # full-depth pass of a plain model with the standard class seed: replayed as one CUDA graph (replay_pass)
            if mask_row is None:
                mask_row = torch.zeros(plan.ops[-1].cout, device=x.device)
                mask_row[fn_key[1]] = 1.0
            out.append(plan.replay_pass(("relevance", tuple(xb.shape), fn_key[2], plan.use_tc), (xb, mask_row), one_pass)[0])
            continue
        logits, saved, _ = plan.forward(xb, keep_from=0, out_index=-1)
        seed = fn(logits) if seed_rows is None else seed_rows(logits, i)
        out.append(plan.backward(seed.contiguous(), saved, stop_after=-1))
    return torch.cat(out, 0)


def lrp_input_relevance(model, input_batch, composite, attr_output_fn: Callable, batch_size: int = 64) -> torch.Tensor:
    """Relevance at the input (compute_relevances, attribute.py:70-108).  The reference attributes the whole batch in
    one pass; here it is cut into minibatches, so a seed function that looks at the batch as a whole (the balanced
    all-classes mask of attribute.py:148-158) is evaluated on the logits of the full batch first.  For a
    ProjectionModel the batch is read as groups of K+1 clones (explainer.py:90-99): identical clones share one forward
    pass and the backward pass above the filter."""
    x = _prep_input(input_batch)
    batch_size = _engine_chunk(x, batch_size)
    with torch.cuda.device(x.device):
        plan = _plan(model, composite, x.device)
        while True:
            if plan.filter_index is None:
                if _seed_is_rowwise(attr_output_fn) or x.size(0) <= batch_size:
                    res = _run_relevance(plan, x, attr_output_fn, batch_size)
                else:
                    logits = torch.cat([plan.forward(x[i:i + batch_size], keep_from=len(plan.ops), out_index=-1)[0]
                                        for i in range(0, x.size(0), batch_size)], 0)
                    seed_full = attr_output_fn(logits)
                    res = _run_relevance(plan, x, None, batch_size, lambda lg, i: seed_full[i:i + lg.size(0)])
            else:
                res = _clone_relevance(plan, x, attr_output_fn, batch_size)
            if not plan.tc_failed():
                break
    return res


def _clone_relevance(plan, x, fn, batch_size):
    Kp1 = plan.filter_K + 1
    N = x.size(0)
    if N % Kp1 != 0:
        raise _L.DRSAError(f"a ProjectionModel attributes groups of num_concepts + 1 = {Kp1} clones; got a batch of {N}")
    B = N // Kp1
    xg = x.view(B, Kp1, *x.shape[1:])
    xu = xg[:, 0].contiguous()
    if bool((xg == xg[:, :1]).all()):
        # the reference's use: every sample repeated K+1 times -> one forward per sample
        bs = max(1, batch_size // Kp1)
        logits = torch.cat([plan.forward(xu[i:i + bs], keep_from=len(plan.ops), out_index=-1)[0] for i in range(0, B, bs)], 0) \
            if not _seed_is_rowwise(fn) else None
        seed_full = None
        if logits is not None:
            seed_full = fn(logits.repeat_interleave(Kp1, dim=0)).view(B, Kp1, -1)
            if not bool((seed_full == seed_full[:, :1]).all()):
                seed_full = False               # clones of one sample are seeded differently: general path below
        if seed_full is not False:
            if seed_full is None:
                return _run_relevance(plan, xu, fn, bs)
            return _run_relevance(plan, xu, None, bs, lambda lg, i: seed_full[i:i + lg.size(0), 0])
    # general case (clones differ): clone index k of every group is attributed on its own and keeps slot k
    logits = torch.cat([plan.forward(x[i:i + batch_size], keep_from=len(plan.ops), out_index=-1)[0]
                        for i in range(0, N, batch_size)], 0)
    seed_all = fn(logits).view(B, Kp1, -1)
    res = torch.empty(B, Kp1, *x.shape[1:], device=x.device)
    for kk in range(Kp1):
        xk = xg[:, kk].contiguous()
        r = _run_relevance(plan, xk, None, max(1, batch_size // Kp1), lambda lg, i: seed_all[i:i + lg.size(0), kk])
        res[:, kk] = r.view(B, Kp1, *x.shape[1:])[:, kk]
    return res.flatten(0, 1)
