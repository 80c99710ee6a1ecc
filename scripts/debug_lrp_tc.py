"""Layer-by-layer comparison of the tensor-core LRP stack against the CUDA-core path (debug aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lrp_ref
from cxai.utils.constants import lrp_name_map_6s
from cxai.xai.explain.rules import NameMapComposite, SequentialMergeBatchNorm
from cxai.xai.explain import lrp_engine
from cxai.xai.explain.attribute import lrp_output_modifier

net = lrp_ref.genre_model(seed=0, last=64, input_size=(32, 64))
x = lrp_ref.synth_logmel(6, 32, 64, 20262).cuda()
comp = NameMapComposite(lrp_name_map_6s(), canonizers=[SequentialMergeBatchNorm()])
plan = lrp_engine._plan(net, comp, x.device)
fn = lrp_output_modifier(3)
print("tc ok:", plan._tc_stack_ok(x), "ops:", [(k, o.kind) for k, o in enumerate(plan.ops)][:12])
res = {}
for use_tc in (True, False):
    plan.use_tc = use_tc
    logits, saved, outs = plan.forward(x, keep_from=0)
    rs = {}
    for stop in list(range(len(plan.ops) - 2, -2, -1)):
        try:
            R = plan.backward(fn(logits).contiguous(), saved, stop_after=stop)
            rs[stop] = R.double().cpu()
        except Exception as e:
            rs[stop] = repr(e)
    res[use_tc] = (logits.double().cpu(), rs)
print("logits diff", float((res[True][0] - res[False][0]).abs().max()))
for stop in sorted(res[True][1].keys(), reverse=True):
    a, b = res[True][1][stop], res[False][1][stop]
    if isinstance(a, str) or isinstance(b, str):
        print(stop, plan.ops[stop].kind if stop >= 0 else "input", "ERR", a if isinstance(a, str) else "", b if isinstance(b, str) else "")
        continue
    if a.shape != b.shape:
        print(stop, "shape", a.shape, b.shape); continue
    rel = float((a - b).norm() / b.norm().clamp(min=1e-30))
    print(f"{stop:3d} {plan.ops[stop].kind if stop >= 0 else 'input':8s} |R_tc|={float(a.norm()):.4e} |R_fp|={float(b.norm()):.4e} rel={rel:.2e}")
print("tc_err", plan._tc_err)
