#!/usr/bin/env python
"""Stress of the pipelined detect path (lean NMS kernel, helper CTAs, priority streams): many steps with alternating and
freshly drawn inputs, every batch's rows compared bitwise with the single-stream plan."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vision_conglomerate_b200 import ops, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
nsets = 6
sets = [[r.to(dev) for r in synth.raw_head_outputs(B, 640, 640, 80, "T" if i % 3 else "TP", 50 + i)] for i in range(nsets)]
ref = []
for raws in sets:
    d = ops.detect(raws, anc, (640, 640), 80, iou_threshold=0.65, score_threshold=0.001, box_allowance=4)
    ref.append((d.pred_boxes.clone(), d.keep_idxs.clone(), d.counts.clone()))
bad = 0
for depth in (3, 4):
    pipe = ops.DetectPipeline([tuple(r.shape) for r in sets[0]], anc, (640, 640), 80, dev, None, 0.65, 0.001, 4, None, depth=depth)
    pending = {}
    for i in range(steps):
        slot = pipe.submitted % depth
        if slot in pending:
            j = pending.pop(slot)
            got = pipe.result(slot)
            ok = torch.equal(got.pred_boxes, ref[j][0]) and torch.equal(got.keep_idxs, ref[j][1]) and torch.equal(got.counts, ref[j][2])
            bad += 0 if ok else 1
        j = (i * 7 + i // 5) % nsets
        pipe.submit(sets[j])
        pending[slot] = j
    for slot, j in pending.items():
        got = pipe.result(slot)
        bad += 0 if (torch.equal(got.pred_boxes, ref[j][0]) and torch.equal(got.keep_idxs, ref[j][1])) else 1
    pipe.join()
    print("depth %d: %d steps, path %d, mismatching batches so far %d" % (depth, steps, pipe.plans[0].params.nms_path, bad), flush=True)
sys.exit(1 if bad else 0)
