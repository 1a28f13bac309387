#!/usr/bin/env python
"""Per-phase cycle breakdown of decode_filter_kernel (thread 0 of every CTA), config 2 dist T."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import _lib, ops, synth
dev = torch.device("cuda", 0)
B, H, W, C = 64, 640, 640, 80
raws = [r.to(dev) for r in synth.raw_head_outputs(B, H, W, C, "T", 7)]
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (H, W), C, dev, None, 0.65, 0.001, 4)
for _ in range(3):
    plan.enqueue(raws); plan.result()
L = _lib.lib()
buf = torch.zeros(1024, 8, dtype=torch.int64, device=dev)
L.bg_profile_decode_cycles(buf.data_ptr())
plan.enqueue(raws); torch.cuda.synchronize()
L.bg_profile_decode_cycles(None)
c = buf.cpu().double()
c = c[c.sum(1) > 0]
names = ["locate+wait", "phase1", "barrier1", "list+barrier", "(unused)", "phase2", "barrier_end", "refill"]
tot = c.sum(1)
print("CTAs %d, cycles per CTA mean %.0f (%.1f us at 1.965 GHz), tiles/CTA %.1f" % (c.shape[0], tot.mean(), tot.mean() / 1965, 64 * 198 / c.shape[0]))
for i, n in enumerate(names):
    print("  %-18s %6.1f%%  %7.0f cycles/tile" % (n, 100 * c[:, i].sum() / tot.sum(), c[:, i].mean() / (64 * 198 / c.shape[0])))
