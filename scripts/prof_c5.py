#!/usr/bin/env python
"""Where the batch-1 latency (config 5) goes: host enqueue, GPU time of the two kernels, result read."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth

dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
frames = [[r.to(dev) for r in synth.raw_head_outputs(1, 640, 640, 80, "TP", 7 + f)] for f in range(32)]
plan = ops.DetectPlan([tuple(r.shape) for r in frames[0]], anc, (640, 640), 80, dev, (720, 1280), 0.35, 0.3, 4, synth.tracked_classes_default())
for f in range(64):
    plan.enqueue(frames[f % 32]); plan.result()
enq, res, tot, gpu = [], [], [], []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for f in range(500):
    t0 = time.perf_counter()
    e0.record()
    plan.enqueue(frames[f % 32])
    e1.record()
    t1 = time.perf_counter()
    r = plan.result()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e6); res.append((t2 - t1) * 1e6); tot.append((t2 - t0) * 1e6)
    gpu.append(e0.elapsed_time(e1) * 1e3)
med = lambda v: sorted(v)[len(v) // 2]
print("p50 us: host enqueue (incl. 2 event records) %.1f, result() %.1f, total %.1f; GPU span of the two kernels %.1f; rows %d"
      % (med(enq), med(res), med(tot), med(gpu), int(r.pred_boxes.shape[0])))
# the same with the rows brought to the HOST by one pinned copy queued behind the kernels (what the reference's
# per-image loop needs, inference_det.py:100-129)
tot3 = []
for f in range(500):
    t0 = time.perf_counter()
    plan.enqueue(frames[f % 32])
    plan.enqueue_host_copy()
    h = plan.result_host()
    tot3.append((time.perf_counter() - t0) * 1e6)
print("enqueue + one pinned copy of [counts | rows] + result_host(): p50 %.1f us, p99 %.1f us, rows on the host %d"
      % (med(tot3), sorted(tot3)[int(len(tot3) * 0.99)], int(h.rows.shape[0])))
tot4 = []
for f in range(500):
    t0 = time.perf_counter()
    plan.enqueue(frames[f % 32])
    r = plan.result()
    tot4.append((time.perf_counter() - t0) * 1e6)
print("enqueue + result() (rows stay on the device): p50 %.1f us, p99 %.1f us" % (med(tot4), sorted(tot4)[int(len(tot4) * 0.99)]))
# host-result plan: the kernels write rows / counts into page-locked host memory and store a flag last; the host polls it
hplan = ops.DetectPlan([tuple(r.shape) for r in frames[0]], anc, (640, 640), 80, dev, (720, 1280), 0.35, 0.3, 4, synth.tracked_classes_default(),
                       host_result=True)
for f in range(64):
    hplan.enqueue(frames[f % 32]); hplan.result_host()
tot5, enq5 = [], []
for f in range(1000):
    t0 = time.perf_counter()
    hplan.enqueue(frames[f % 32])
    t1 = time.perf_counter()
    h = hplan.result_host()
    tot5.append((time.perf_counter() - t0) * 1e6); enq5.append((t1 - t0) * 1e6)
print("host-result plan (no copy, no sync; poll the flag): p50 %.1f us, p99 %.1f us (enqueue %.1f), rows on the host %d"
      % (med(tot5), sorted(tot5)[int(len(tot5) * 0.99)], med(enq5), int(h.rows.shape[0])))
# CUDA graph replay of the same two launches from a static input
static = [t.clone() for t in frames[0]]
plan.enqueue(static); plan.result()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    plan.enqueue(static)
torch.cuda.synchronize()
tot2 = []
for f in range(500):
    t0 = time.perf_counter()
    for s, src in zip(static, frames[f % 32]):
        s.copy_(src, non_blocking=True)
    g.replay()
    r = plan.result()
    tot2.append((time.perf_counter() - t0) * 1e6)
print("graph replay (3 d2d input copies + replay + result): p50 %.1f us, p99 %.1f us, rows %d" % (med(tot2), sorted(tot2)[int(len(tot2) * 0.99)], int(r.pred_boxes.shape[0])))
# stage stamps of the per-image NMS kernel for one frame (%globaltimer hook)
from vision_conglomerate_b200 import _lib
L = _lib.lib()
ns = int(L.bg_profile_stamps_per_image())
st = torch.zeros(1, ns, dtype=torch.int64, device=dev)
L.bg_profile_stamps(st.data_ptr())
plan.enqueue(frames[0]); torch.cuda.synchronize()
L.bg_profile_stamps(None)
s = st.cpu()[0].tolist()
names = ["load_slots", "grid_bucket", "pair_tests", "resolve", "sort", "rank+lookback", "write_rows"]
print("NMS stages (us):", {n: round((s[i + 1] - s[i]) / 1e3, 2) for i, n in enumerate(names)}, "candidates", int(plan.result().candidates[0]))
