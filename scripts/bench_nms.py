#!/usr/bin/env python
"""Timing of the general (segmented) NMS engine: bg_batched_nms over all candidates of a batch (the call the
unmodified inference script makes, B4) and the general path of bg_detect when every candidate survives (c2R)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth

dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
out = {}


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for dist in ("T", "R"):
    B = 64
    raws = [r.to(dev) for r in synth.raw_head_outputs(B, 640, 640, 80, dist, 7)]
    plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (640, 640), 80, dev, None, 0.6, 0.0 if dist == "R" else 0.25, 4, None,
                          nms_path="general")
    plan.enqueue(raws)
    r = plan.result()
    ms = timed(lambda: plan.enqueue(raws), 5)
    out["detect_general_" + dist] = {"ms": ms, "kept": int(r.pred_boxes.shape[0]), "cand_per_img": float(r.candidates.float().mean()),
                                     "mask_bytes": plan.mask_bytes}
    # B4: every candidate of the batch through batched_nms, groups = images
    preds = torch.cat([ops.decode_scale(rw, a, (640, 640), inference=True).reshape(B, -1, 85) for rw, a in zip(raws, anc)], 1)
    boxes = preds[..., :4].reshape(-1, 4)
    xyxy = torch.cat([boxes[:, :2] - boxes[:, 2:] / 2, boxes[:, :2] + boxes[:, 2:] / 2], 1).contiguous()
    sc = (preds[..., 4] * preds[..., 5:].max(-1).values).reshape(-1).contiguous()
    idx = torch.arange(B, device=dev).repeat_interleave(preds.shape[1])
    k = ops.batched_nms(xyxy, sc, idx, 0.6)
    ms = timed(lambda: ops.batched_nms(xyxy, sc, idx, 0.6), 3)
    out["batched_nms_all_" + dist] = {"ms": ms, "n": int(sc.numel()), "kept": int(k.numel())}
print(json.dumps(out))
