#!/usr/bin/env python
"""One bg_batched_nms call over every candidate of a 64-image batch (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth

dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
dist = sys.argv[1] if len(sys.argv) > 1 else "R"
B = 64
raws = [r.to(dev) for r in synth.raw_head_outputs(B, 640, 640, 80, dist[-1], 7)]
if dist.startswith("det"):
    d = dist[3:]
    plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (640, 640), 80, dev, None, 0.6, 0.0 if d == "R" else 0.25, 4, None,
                          nms_path="general")
    for _ in range(3):
        plan.enqueue(raws)
        r = plan.result()
    print(dist, int(r.pred_boxes.shape[0]))
    sys.exit(0)
preds = torch.cat([ops.decode_scale(rw, a, (640, 640), inference=True).reshape(B, -1, 85) for rw, a in zip(raws, anc)], 1)
boxes = preds[..., :4].reshape(-1, 4)
xyxy = torch.cat([boxes[:, :2] - boxes[:, 2:] / 2, boxes[:, :2] + boxes[:, 2:] / 2], 1).contiguous()
sc = (preds[..., 4] * preds[..., 5:].max(-1).values).reshape(-1).contiguous()
idx = torch.arange(B, device=dev).repeat_interleave(preds.shape[1])
for _ in range(3):
    k = ops.batched_nms(xyxy, sc, idx, 0.6)
torch.cuda.synchronize()
print(dist, int(k.numel()))
