"""CPU baseline of STAGE 1 run by the reference's own code: ``python oracle/ref_baseline.py lrp <samples>`` prints one JSON
object.  Runs in its own process because the reference package is also called ``cxai``: ``oracle/_ref`` (byte copy made by
oracle/make_ref.py) goes in front of ``sys.path``, ``oracle/mini_zennit`` is registered as ``zennit`` (the real package is
unobtainable here), and ``cxai.xai.drsa.preprocessing.get_intermediate`` (preprocessing.py:106-176) is timed on the
BASELINE cfg 2 CNN built by the reference's ``VGGType`` constructor.  TEST / BASELINE INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def main():
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ref_root = os.path.join(HERE, "_ref")
    if not os.path.isfile(os.path.join(ref_root, "cxai/xai/drsa/preprocessing.py")):
        print(json.dumps({"unavailable": "oracle/_ref missing (run oracle/make_ref.py in the build container)"}))
        return
    sys.path = [ref_root] + [p for p in sys.path if os.path.abspath(p or ".") != REPO] + [REPO]
    import torch
    from oracle import mini_zennit, synth
    mini_zennit.install()
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    import cxai
    assert os.path.abspath(cxai.__file__).startswith(ref_root), cxai.__file__
    from cxai.model.create_model import VGGType
    from cxai.xai.drsa.preprocessing import get_intermediate
    from zennit.rules import Epsilon, Gamma, WSquare
    from zennit.composites import NameMapComposite
    from zennit.canonizers import SequentialMergeBatchNorm

    class FlatVGG(VGGType):                                  # create_model.py:95 hard-codes 2048 flat features
        def forward(self, x):
            x = self.features(x)
            return self.classifier(x.view(x.size(0), -1))

    g, st = 0.3, 1e-7                                        # getdrsadata.py:81-108
    name_map = [(['features.0'], WSquare(stabilizer=st))] + \
        [([f'features.{i}'], Gamma(gamma=gm, stabilizer=st)) for i, gm in
         ((3, g), (7, g), (10, g), (14, g / 2), (17, g / 2), (21, g / 2), (24, g / 2), (28, g / 4), (31, g / 4))] + \
        [([f'classifier.{i}'], Epsilon(epsilon=st)) for i in (0, 4, 8)]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    net = synth.build_model(FlatVGG, "cfg2", seed=0, bn_seed=None)
    x = synth.synth_logmel(n, 128, 256, 20262)
    comp = lambda: NameMapComposite(name_map, canonizers=[SequentialMergeBatchNorm()])
    import contextlib, io
    with contextlib.redirect_stderr(io.StringIO()):          # tqdm bars of the reference
        get_intermediate(net, x[:1], comp(), net.features[33], 0)          # warm-up
        t0 = time.perf_counter()
        a, R = get_intermediate(net, x, comp(), net.features[33], 0)
        t = time.perf_counter() - t0
    P = a.shape[-1] * a.shape[-2]
    print(json.dumps({"value": n * P / t, "unit": "vectors/s", "cores": threads, "kind": "reference",
                      "sample": f"{n} samples of 128x256 through the reference's own get_intermediate (preprocessing.py:106-176, "
                                f"byte copy in oracle/_ref) on the mini-zennit restatement of zennit 0.5.1, torch {torch.__version__} "
                                f"CPU fp32, {threads} threads, {P} positions each", "sample_s": t}))


if __name__ == "__main__":
    main()
