#!/usr/bin/env python
"""Profiling driver for the general NMS engine: config 4 decode+NMS (B=32, 1280^2, dist T, ~6.8k survivors/img)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth
dev = torch.device("cuda", 0)
B, S = 32, 1280
raws = [r.to(dev) for r in synth.raw_head_outputs(B, S, S, 80, "T", 7)]
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (S, S), 80, dev, None, 0.65, 0.001, 4)
for _ in range(3):
    plan.enqueue(raws); r = plan.result()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    plan.enqueue(raws)
e1.record(); torch.cuda.synchronize()
print("c4 detect: %.3f ms/batch, path %s, kept %d" % (e0.elapsed_time(e1) / 5, "general" if plan.params.nms_path == 1 else "per-image", r.pred_boxes.shape[0]))
