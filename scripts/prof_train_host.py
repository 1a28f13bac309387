#!/usr/bin/env python
"""Host-side cost of one training step of the loss (enqueue only, no sync) against the GPU time, at a small batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
t = synth.targets(B, 100, 80, 0).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
preds = [torch.randn(B, ny, nx, 3, 85, generator=g, device=dev).requires_grad_(True) for ny, nx in synth.fmap_shapes(640, 640)]
for _ in range(5):
    for p in preds: p.grad = None
    ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False)[0].backward()
torch.cuda.synchronize()
N = 200
tf = tb = 0.0
t0 = time.perf_counter()
for _ in range(N):
    for p in preds: p.grad = None
    a = time.perf_counter()
    loss = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False)[0]
    b = time.perf_counter()
    loss.backward()
    c = time.perf_counter()
    tf += b - a; tb += c - b
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("B=%d: host enqueue per step: forward %.1f us, backward %.1f us, loop %.1f us; wall incl. final sync %.1f us/step"
      % (B, tf / N * 1e6, tb / N * 1e6, (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
