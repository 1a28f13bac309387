#!/usr/bin/env python
"""Pipelined decode+NMS step (ops.DetectPipeline) at config 2 with the lean / the 1024-thread per-image NMS kernel and
several depths: ms per batch over K steps (CUDA events), results checked against the single-stream plan."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vision_conglomerate_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depths", default="2,4,6")
ap.add_argument("--steps", default="20,100")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=640)
ap.add_argument("--lean", default="0,1")
ap.add_argument("--stages", type=int, default=1)
ap.add_argument("--prio", default="1")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, C = args.batch, args.size, 80
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
sets = [[r.to(dev) for r in synth.raw_head_outputs(B, S, S, C, "T", 7 + 100 * i)] for i in range(2)]
shapes = [tuple(r.shape) for r in sets[0]]
ref = []
for raws in sets:
    d = ops.detect(raws, anc, (S, S), C, iou_threshold=0.65, score_threshold=0.001, box_allowance=4)
    ref.append((d.pred_boxes.clone(), d.keep_idxs.clone()))
print("survivors/img %.0f max %d, kept %d" % (float(d.candidates.float().mean()), int(d.candidates.max()), ref[1][0].shape[0]))
for lean in [bool(int(x)) for x in args.lean.split(",")]:
    ops.LEAN_NMS = lean
    for prio, depth in [(bool(int(p)), int(x)) for p in args.prio.split(",") for x in args.depths.split(",")]:
        pipe = ops.DetectPipeline(shapes, anc, (S, S), C, dev, None, 0.65, 0.001, 4, None, depth=depth, nms_priority=prio)
        for it in range(3):
            for _ in range(depth):
                pipe.submit(sets[it & 1])
            for s in range(depth):
                pipe.result(s)
        for i in range(depth):
            pipe.submit(sets[i & 1])
        for s in range(depth):
            got = pipe.result(s)
            assert torch.equal(got.pred_boxes, ref[s & 1][0]) and torch.equal(got.keep_idxs, ref[s & 1][1]), (lean, depth, s)
        pipe.join()
        torch.cuda.synchronize()
        out = []
        for K in [int(x) for x in args.steps.split(",")]:
            best, tot = 1e9, 0.0
            for _r in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for i in range(K):
                    pipe.submit(sets[i & 1])
                pipe.join()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
                tot += e0.elapsed_time(e1) / K / 3
            out.append("K=%d: best %.1f mean %.1f us" % (K, best * 1e3, tot * 1e3))
        print("lean=%d prio=%d path=%d depth=%d  %s" % (lean, prio, pipe.plans[0].params.nms_path, depth, "  ".join(out)), flush=True)

if not args.stages:
    sys.exit(0)
# stage breakdown of the per-image kernels running alone (one plan, one stream)
from vision_conglomerate_b200 import _lib  # noqa: E402
L = _lib.lib()
names = ["load_slots", "grid_bucket", "pair_tests", "resolve", "sort", "rank+lookback", "write_rows"]
for path in ("per_image_lean",):
    plan = ops.DetectPlan(shapes, anc, (S, S), C, dev, None, 0.65, 0.001, 4, None, "image", 0, path)
    for _ in range(3):
        plan.enqueue(sets[0]); plan.result()
    ns = int(L.bg_profile_stamps_per_image())
    stamps = torch.zeros(B, ns, dtype=torch.int64, device=dev)
    L.bg_profile_stamps(stamps.data_ptr())
    plan.enqueue(sets[0])
    torch.cuda.synchronize()
    L.bg_profile_stamps(None)
    st = stamps.cpu().double()
    d = {n: round(float((st[:, i + 1] - st[:, i]).mean()) / 1e3, 1) for i, n in enumerate(names)}
    print(path, d, "span %.1f mean %.1f" % (float(st[:, 7].max() - st[:, 0].min()) / 1e3, float((st[:, 7] - st[:, 0]).mean()) / 1e3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(20):
        plan.enqueue(sets[i & 1])
    e1.record()
    torch.cuda.synchronize()
    print("   single-stream batch latency %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
