#!/usr/bin/env python
"""Diagnose run-to-run variation of the training step: per-iteration forward / backward times (CUDA events),
allocator activity (device allocations during the timed loop) and host enqueue time."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vision_conglomerate_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--form", default="raw")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--prelude-fwd", type=int, default=0, help="forward-only calls (graph kept, no backward) before the timed loop, like prof_train.py")
ap.add_argument("--nosync", action="store_true", help="enqueue all iterations back to back (no per-iteration synchronize)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, C = a.batch, 640, 80
t = synth.targets(B, 100, C, 0).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
preds = [torch.randn(B, ny, nx, 3, 5 + C, generator=g, device=dev).requires_grad_(True) for ny, nx in synth.fmap_shapes(S, S)]
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
for _ in range(3):
    for p in preds:
        p.grad = None
    loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False, input_form=a.form)
    loss.backward()
torch.cuda.synchronize()
for _ in range(a.prelude_fwd):
    for p in preds:
        p.grad = None
    loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False, input_form=a.form)
st0 = torch.cuda.memory_stats()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.iters)]
host = []
for i in range(a.iters):
    for p in preds:
        p.grad = None
    h0 = time.perf_counter()
    ev[i][0].record()
    loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False, input_form=a.form)
    ev[i][1].record()
    h1 = time.perf_counter()
    loss.backward()
    ev[i][2].record()
    h2 = time.perf_counter()
    host.append(((h1 - h0) * 1e3, (h2 - h1) * 1e3))
    if not a.nosync:
        torch.cuda.synchronize()
torch.cuda.synchronize()
st1 = torch.cuda.memory_stats()
print('total GPU span %.3f ms/iter, host enqueue %.3f ms/iter, loadavg %s, cpus %d' % (ev[0][0].elapsed_time(ev[-1][2]) / a.iters, sum(x + y for x, y in host) / a.iters, os.getloadavg(), len(os.sched_getaffinity(0))))
print("form=%s B=%d env PDL=%s" % (a.form, B, os.environ.get("BG_PDL", "1")))
for i in range(a.iters):
    print("  it %2d: fwd %.3f ms  bwd %.3f ms | host fwd %.3f bwd %.3f" % (i, ev[i][0].elapsed_time(ev[i][1]), ev[i][1].elapsed_time(ev[i][2]),
                                                                        host[i][0], host[i][1]))
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated", "segment.all.allocated"):
    print("  %s: +%d" % (k, st1.get(k, 0) - st0.get(k, 0)))
print("  reserved %.2f GB, allocated %.2f GB" % (torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_allocated() / 1e9))
