#!/usr/bin/env python
"""Prototype: consecutive batches alternate between two streams (own plan, scratch and outputs each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth

dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
B = 64
raws = [r.to(dev) for r in synth.raw_head_outputs(B, 640, 640, 80, "T", 7)]
raws2 = [r.to(dev) for r in synth.raw_head_outputs(B, 640, 640, 80, "T", 8)]
for nstreams in (1, 3, 4, 6):
    plans = []
    for i in range(nstreams):
        pl = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (640, 640), 80, dev, None, 0.65, 0.001, 4, None)
        pl.ws_tag = "detect%d" % i
        plans.append(pl)
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    inputs = [raws, raws2, raws, raws2, raws, raws2]
    def run(K):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for k in range(K):
            with torch.cuda.stream(streams[k % nstreams]):
                plans[k % nstreams].enqueue(inputs[k % nstreams])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K
    run(6)
    ms = run(60)
    rows = [int(p.result().pred_boxes.shape[0]) for p in plans]
    print(nstreams, "streams: %.2f us/step" % (ms * 1e3), rows)
