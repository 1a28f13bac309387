#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configurations other than the headline (configs[1]):
  c2R  decode+NMS, every candidate survives (general NMS engine, ALU-bound stress)
  c3   training-step assignment + loss fwd+bwd, B=256 (one GPU's view of the whole batch) and the B=32 shard
  c4   1280x1280, 300 gt/img, B=32: assignment+loss, decode+NMS on dist T
  c5   batch-1 latency at 640^2, IoU 0.35 / score 0.3, tracked classes, 1000 frames (p50/p99)
  tv   torchvision-CUDA batched_nms on the same decoded candidates (the existing sm_100 kernel)
Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_conglomerate_b200 import ops, synth

dev = torch.device("cuda", 0)
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
out = {}


def timed(fn, iters, warm=3):
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


def detect_case(name, B, S, dist, iou, thr, iters=10, tracked=None):
    raws = [r.to(dev) for r in synth.raw_head_outputs(B, S, S, 80, dist, 7)]
    plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (S, S), 80, dev, None, iou, thr, 4, tracked)
    plan.enqueue(raws)
    r = plan.result()
    ms = timed(lambda: plan.enqueue(raws), iters)
    # the same configuration with four batches in flight (ops.DetectPipeline)
    pipe = ops.DetectPipeline([tuple(r.shape) for r in raws], anc, (S, S), 80, dev, None, iou, thr, 4, tracked, depth=4)
    for _ in range(2):
        for _d in range(4):
            pipe.submit(raws)
        for d in range(4):
            pipe.result(d)
    pipe.join()

    def piped():
        for _d in range(8):
            pipe.submit(raws)
        pipe.join()
    ms_p = timed(piped, max(2, iters // 2)) / 8
    out[name] = {"ms_per_batch": ms, "img_per_s": B / ms * 1e3, "ms_per_batch_pipelined": ms_p, "img_per_s_pipelined": B / ms_p * 1e3,
                 "hbm_frac_pipelined": plan.input_bytes / (ms_p * 1e-3) / 1e9 / 6552.6,
                 "batch": B, "kept_rows": int(r.pred_boxes.shape[0]),
                 "survivors_per_image": float(r.candidates.float().mean()), "nms_path": "general" if plan.params.nms_path == 1 else "per-image",
                 "hbm_frac_of_measured": plan.input_bytes / (ms * 1e-3) / 1e9 / 6552.6}
    return raws, plan


def train_case(name, B, S, G, iters=10):
    t = synth.targets(B, G, 80, 0).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    preds = [torch.randn(B, ny, nx, 3, 85, generator=g, device=dev).requires_grad_(True) for ny, nx in synth.fmap_shapes(S, S)]

    def fwd():
        return ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False)[0]

    def step():
        for p in preds:
            p.grad = None
        fwd().backward()
    f = timed(fwd, iters)
    b = timed(step, iters)
    asg = timed(lambda: [ops.build_target_by_scale(t, (ny, nx), a) for (ny, nx), a in zip(synth.fmap_shapes(S, S), anc)], iters)
    alg = sum(p.numel() for p in preds) * 4 / B * 1.18  # ~ grad write + strided objectness plane + matched rows
    out[name] = {"fwd_ms": f, "fwd_bwd_ms": b, "img_per_s": B / b * 1e3, "assign_only_ms_3_scales_with_sync": asg, "batch": B, "gt_per_img": G,
                 "hbm_frac_of_measured": B * alg / (b * 1e-3) / 1e9 / 6552.6}


detect_case("c2R_all_candidates_survive", 64, 640, "R", 0.65, 0.001, iters=3)
train_case("c3_B256", 256, 640, 100)
train_case("c3_shard_B32", 32, 640, 100)
train_case("c4_train_B32_1280", 32, 1280, 300)
detect_case("c4_detect_T_B32_1280", 32, 1280, "T", 0.65, 0.001)

# c5: batch-1 latency, per-frame inputs, host-visible result every frame
lat = []
frames = [[r.to(dev) for r in synth.raw_head_outputs(1, 640, 640, 80, "TP", 7 + f)] for f in range(32)]
plan = ops.DetectPlan([tuple(r.shape) for r in frames[0]], anc, (640, 640), 80, dev, (720, 1280), 0.35, 0.3, 4, synth.tracked_classes_default())
for f in range(1032):
    t0 = time.perf_counter()
    plan.enqueue(frames[f % 32])
    r = plan.result()
    lat.append((time.perf_counter() - t0) * 1e6)
lat = sorted(lat[32:])
out["c5_batch1_latency_us"] = {"p50": lat[len(lat) // 2], "p99": lat[int(len(lat) * 0.99)], "mean": sum(lat) / len(lat), "frames": len(lat),
                               "rows_last_frame": int(r.pred_boxes.shape[0]), "note": "enqueue + one D2H count read + sync per frame, wall clock"}

# the same with the rows on the HOST every frame: the kernels write into page-locked host memory and store a flag last
# (DetectPlan(host_result=True), bg_detect_params.host_flag); no device->host copy, no stream synchronisation
hplan = ops.DetectPlan([tuple(r.shape) for r in frames[0]], anc, (640, 640), 80, dev, (720, 1280), 0.35, 0.3, 4, synth.tracked_classes_default(),
                       host_result=True)
lat = []
for f in range(1032):
    t0 = time.perf_counter()
    hplan.enqueue(frames[f % 32])
    h = hplan.result_host()
    lat.append((time.perf_counter() - t0) * 1e6)
lat = sorted(lat[32:])
out["c5_batch1_latency_host_rows_us"] = {"p50": lat[len(lat) // 2], "p99": lat[int(len(lat) * 0.99)], "mean": sum(lat) / len(lat), "frames": len(lat),
                                         "rows_last_frame": int(h.rows.shape[0]),
                                         "note": "enqueue + poll of the flag the last kernel stores in page-locked host memory; rows [k, 6] on the host, wall clock"}

# torchvision-CUDA on the same candidates: decode with our kernel, then its batched_nms over ALL candidates (the reference path)
try:
    import torchvision
    B, S = 64, 640
    raws = [r.to(dev) for r in synth.raw_head_outputs(B, S, S, 80, "T", 7)]
    dec = torch.cat([ops.decode_scale(r, a, (S, S), True).reshape(B, -1, 85) for r, a in zip(raws, anc)], 1)
    p = dec.reshape(-1, 85)
    scores = torch.sigmoid(p[:, 1:81]).max(1)[0] * torch.sigmoid(p[:, 0])
    xywh = p[:, 81:85].clone(); xywh[:, 2:] += 4
    xyxy = torch.cat([xywh[:, :2] - xywh[:, 2:] / 2, xywh[:, :2] - xywh[:, 2:] / 2 + xywh[:, 2:]], 1).contiguous()
    idxs = torch.arange(B, device=dev).repeat_interleave(p.shape[0] // B)
    sel = scores > 0.001
    bs, ss, ii = xyxy[sel].contiguous(), scores[sel].contiguous(), idxs[sel].contiguous()
    ms_f = timed(lambda: torchvision.ops.batched_nms(bs, ss, ii, 0.65), 5, warm=2)
    ms_ours = timed(lambda: ops.batched_nms(bs, ss, ii, 0.65), 5, warm=2)
    out["batched_nms_prefiltered_109k_boxes_64_groups"] = {"torchvision_cuda_ms": ms_f, "bg_batched_nms_ms": ms_ours}
    ms_all = timed(lambda: torchvision.ops.batched_nms(xyxy, scores, idxs, 0.65), 1, warm=1)
    ms_all_ours = timed(lambda: ops.batched_nms(xyxy, scores, idxs, 0.65), 1, warm=1)
    out["batched_nms_all_1p6M_candidates_64_groups"] = {"torchvision_cuda_ms": ms_all, "bg_batched_nms_ms": ms_all_ours,
                                                        "note": "what inference_det.py:77 does: NMS over every candidate, threshold after"}
except Exception as e:  # noqa: BLE001
    out["torchvision_error"] = repr(e)
print(json.dumps(out))
