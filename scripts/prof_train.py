#!/usr/bin/env python
"""Profiling driver for the training side (config 3 shard): N iterations of the fused detection loss
forward + backward through the public operator.  Run plain for CUDA-event timings, or under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vision_conglomerate_b200 import ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--gt", type=int, default=100)
ap.add_argument("--size", type=int, default=640)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--form", default="decoded", choices=["decoded", "raw", "split"], help="input form of ops.detection_loss")
ap.add_argument("--l2-fetch", type=int, default=0, help="experiment: cudaLimitMaxL2FetchGranularity in bytes (32/64/128)")
a = ap.parse_args()
if a.l2_fetch:
    import ctypes
    torch.cuda.init()
    rt = ctypes.CDLL("libcudart.so.12")
    cur = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(cur), 5)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(a.l2_fetch))  # cudaLimitMaxL2FetchGranularity = 0x05
    new = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(new), 5)
    print("L2 fetch granularity: was %d, set rc=%d, now %d" % (cur.value, rc, new.value))
dev = torch.device("cuda", 0)
B, S, C = a.batch, a.size, 80
t = synth.targets(B, a.gt, C, 0).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
preds = [torch.randn(B, ny, nx, 3, 5 + C, generator=g, device=dev).requires_grad_(True) for ny, nx in synth.fmap_shapes(S, S)]
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
if a.form == "split":
    preds = [tuple(y.contiguous().requires_grad_(True) for y in (x.detach()[..., 0], x.detach()[..., 1:1 + C], x.detach()[..., 1 + C:]))
             for x in preds]
leaves = [q for p in preds for q in (p if isinstance(p, tuple) else (p,))]


def step():
    for p in leaves:
        p.grad = None
    loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False, input_form=a.form)
    loss.backward()
    return loss


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(a.iters):
    for p in leaves:
        p.grad = None
    loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False, input_form=a.form)
e[1].record()
for _ in range(a.iters):
    step()
e[2].record()
torch.cuda.synchronize()
fwd = e[0].elapsed_time(e[1]) / a.iters
both = e[1].elapsed_time(e[2]) / a.iters
print("train form=%s B=%d gt=%d S=%d: fwd %.3f ms, fwd+bwd %.3f ms (%.0f img/s), loss %.6f" % (a.form, B, a.gt, S, fwd, both, B / both * 1e3, float(loss)))

if a.form != "decoded" and os.environ.get("BG_GRAPH", "1") == "1":   # the same step replayed from a CUDA graph (ops.LossStepGraph)
    cfg = dict(synth.LOSS_CONFIG, num_classes=C)
    clones = [tuple(q.detach().clone().requires_grad_(True) for q in p) if isinstance(p, tuple) else p.detach().clone().requires_grad_(True)
              for p in preds]
    gs = ops.LossStepGraph(clones, t, anc, cfg, input_form=a.form)
    for _ in range(3):
        gs.replay()
    best = 1e9
    for _r in range(3):
        x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x.record()
        for _ in range(20):
            gs.replay()
        y.record()
        torch.cuda.synchronize()
        best = min(best, x.elapsed_time(y) / 20)
    print("train form=%s B=%d graph replay: %.4f ms per step (%.0f img/s)" % (a.form, B, best, B / best * 1e3))
