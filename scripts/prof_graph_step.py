#!/usr/bin/env python
"""Graph-replayed loss step (ops.LossStepGraph) at several shard sizes, with the split form's gradient planes cleared
inside the backward or on a second stream next to the forward (ops.PRECLEAR_SPLIT_GRADS)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vision_conglomerate_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
C = 80
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
cfg = dict(synth.LOSS_CONFIG, num_classes=C)
for B in (32, 64, 256):
    t = synth.targets(B, 100, C, 0).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    raw = [torch.randn(B, ny, nx, 3, 5 + C, generator=g, device=dev) for ny, nx in synth.fmap_shapes(640, 640)]
    cells = [x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] for x in raw]
    for form in ("split", "raw"):
        for pre in (("priority", True, False) if form == "split" else (False,)):
            ops.PRECLEAR_SPLIT_GRADS = pre
            if form == "split":
                inp = [tuple(y.contiguous().requires_grad_(True) for y in (x[..., 0], x[..., 1:1 + C], x[..., 1 + C:])) for x in raw]
            else:
                inp = [x.clone().requires_grad_(True) for x in raw]
            gs = ops.LossStepGraph(inp, t, anc, cfg, input_form=form, cells=cells)
            for _ in range(5):
                gs.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _r in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(50):
                    gs.replay()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / 50)
            print("graph step B=%d form=%s preclear=%s: %.4f ms (%.0f img/s), loss %.6f" % (B, form, pre, best, B / best * 1e3, float(gs.loss)))
            del gs, inp
