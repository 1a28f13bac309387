#!/usr/bin/env python
"""bench.py -- headline benchmark of the box-geometry hot path on B200.

    python bench.py --gpus 1 --steps 20 --warmup 3             # our arm (CUDA, sm_100a)
    python bench.py --impl reference --steps 2 --warmup 1      # reference arm: CPU oracle port, all host threads

Metric (BASELINE.json): images/s of inference decode + NMS at 640^2, batch 64 (configs[1]), plus the
fraction of the measured HBM peak sustained by the dominant kernel.  One "step" = one pass of the fused
decode -> score filter -> per-image NMS -> row assembly over one synthetic batch.

  value      device-resident inputs, K steps enqueued through the C ABI, timed with CUDA events;
  e2e        the public API (`ops.DetectPlan`) with pinned-host inputs: H2D copy of the three head
             tensors + kernels + D2H read of the result rows, every step inside the timed region;
  roofline   the decode+filter kernel timed live with CUDA events recorded around it on its stream;
  cpu_baseline  the CPU oracle (port of the reference path) on a bounded sample of the same workload.

N > 1 (torchrun): images are sharded across ranks (every rank runs its own batch of 64; no data-path
collective), barrier + synchronize on both sides, max over ranks, whole-job img/s = N*B*K / t_max.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="c2T", B=64, H=640, W=640, C=80, dist="T", seed=7, iou=0.65, score=0.001, allow=4)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel, per launch, from the committed
    `ncu --set full` capture (profiles/traffic.json); None until a capture of the current kernel exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["decode_filter_kernel"]["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_run(sample_images, steps=1, warmup=0):
    """Times the CPU oracle (port of the reference's decode + post_process path) on `sample_images` images
    of the headline workload.  Returns (img/s, threads, seconds per step)."""
    from oracle import oracle as O
    from vision_conglomerate_b200 import synth
    w = WORKLOAD
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU baseline is meant to use every host core
    threads = os.cpu_count() or 1
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    except OSError:
        threads = int(os.environ.get("OMP_NUM_THREADS", threads))
    raws = synth.raw_head_outputs(sample_images, w["H"], w["W"], w["C"], w["dist"], w["seed"])
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        preds = O.decode_inference(raws, anc, w["H"], w["W"], None)
        out = O.post_process(preds, w["iou"], w["score"], w["allow"], None)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return sample_images / t, int(threads), t, int(out["keep"].shape[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = min(WORKLOAD["B"], max(8, cores))
    ips, threads, t, kept = cpu_oracle_run(sample, steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "images/s (decode+NMS)", "value": ips, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t * WORKLOAD["B"] / sample,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(),
        "cpu_baseline": {"value": ips, "unit": "img/s", "cores": threads, "kind": "port",
                         "sample": "%d of %d images per step; CPU oracle (oracle/boxgeom_oracle.c): decode, score, "
                                   "torchvision-CPU-equivalent greedy NMS over all 25,200 candidates per image, "
                                   "one OpenMP thread per image" % (sample, WORKLOAD["B"])},
        "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _config():
    w = WORKLOAD
    return {"workload": "configs[1] inference decode+NMS: batch %d at %dx%d, %d classes, 25,200 candidates/img, "
                        "conf %.3f, IoU %.2f, box_allowance %d, trained-like logits (dist T, ~1,700 survivors/img)"
                        % (w["B"], w["H"], w["W"], w["C"], w["score"], w["iou"], w["allow"]),
            "per_gpu_batch": w["B"], "parallelism": "image-sharded replicas, no data-path collective",
            "l2": "inputs (548 MB per step) exceed the 126 MB L2; no explicit flush",
            "pipelining": "consecutive batches in flight on separate CUDA streams (ops.DetectPipeline, --depth); "
                          "every step is a full decode+NMS of its batch into its own output buffers"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vision_conglomerate_b200 import _lib, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    devc = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=devc)
    w = WORKLOAD
    B, H, W, C = w["B"], w["H"], w["W"], w["C"]
    # every rank owns its own image shard (different seed -> different images)
    raws_h = [r.pin_memory() for r in synth.raw_head_outputs(B, H, W, C, w["dist"], w["seed"] + rank)]
    raws_d = [r.to(devc) for r in raws_h]
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    plan = ops.DetectPlan([tuple(r.shape) for r in raws_d], anc, (H, W), C, devc, None, w["iou"], w["score"],
                          w["allow"], None, "image", args.variant)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the headline runs consecutive batches through ops.DetectPipeline: `depth` batches in flight on as many CUDA
    # streams (own scratch and outputs each), so one batch's NMS tail overlaps the next batch's HBM-bound decode
    depth = max(1, args.depth)
    pipe = ops.DetectPipeline([tuple(r.shape) for r in raws_d], anc, (H, W), C, devc, None, w["iou"], w["score"],
                              w["allow"], None, "image", args.variant, depth=depth)

    # warm-up (also settles the workspace / mask budget), every slot of the pipeline and the single-stream plan
    for _ in range(max(args.warmup, 3)):
        plan.enqueue(raws_d)
        det = plan.result()
        for _d in range(depth):
            pipe.submit(raws_d)
        for d in range(depth):
            det_p = pipe.result(d)
    pipe.join()
    kept_rows = int(det.pred_boxes.shape[0])
    survivors = float(det.candidates.float().mean())
    if int(det_p.pred_boxes.shape[0]) != kept_rows or not torch.equal(det_p.pred_boxes, det.pred_boxes):
        raise RuntimeError("pipelined and single-stream results differ")

    # ---- value: device-resident inputs, K steps back to back, CUDA events -----------------------
    K = args.steps
    ev_k = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for a, b in ev_k:  # torch creates the cudaEvent lazily: record once so the handle exists
        a.record()
        b.record()
    barrier()
    launches0 = _lib.launch_count()
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            pipe.submit(raws_d)
        pipe.join()
        e1.record()
        launches = _lib.launch_count() - launches0  # kernels of ours enqueued inside the timed region
        barrier()
        # the same K steps on one stream, one batch at a time: the latency view of a step
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for i in range(K):
            plan.enqueue(raws_d)
        l1.record()
        barrier()
        single_ms = l0.elapsed_time(l1) / K
        # second pass, untimed as a whole: the same K steps with CUDA events recorded immediately around the
        # decode+filter kernel on its stream (the roofline numerator's duration)
        for i in range(K):
            L.bg_profile_events(ev_k[i][0].cuda_event, ev_k[i][1].cuda_event)
            plan.enqueue(raws_d)
        barrier()
        L.bg_profile_events(None, None)
        # the timed region lasts milliseconds; keep the same load running ~0.5 s more (untimed) so the
        # 100 ms nvidia-smi sampler sees clocks under this load
        t_end = time.time() + 0.5
        while time.time() < t_end:
            plan.enqueue(raws_d)
            torch.cuda.synchronize()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # stage breakdown of the per-image NMS kernel (one extra untimed step with the %globaltimer hook armed)
    nms_stages = None
    try:
        ns = int(L.bg_profile_stamps_per_image())
        stamps = torch.zeros(B, ns, dtype=torch.int64, device=devc)
        L.bg_profile_stamps(stamps.data_ptr())
        plan.enqueue(raws_d)
        torch.cuda.synchronize()
        L.bg_profile_stamps(None)
        st = stamps.cpu().double()
        if float(st[:, 7].min()) > 0:
            names = ["load_slots", "grid_bucket", "pair_tests", "resolve", "sort", "rank+lookback", "write_rows"]
            nms_stages = {n: float((st[:, i + 1] - st[:, i]).mean()) / 1e3 for i, n in enumerate(names)}
            nms_stages["kernel_span_us"] = float(st[:, 7].max() - st[:, 0].min()) / 1e3
            nms_stages["per_image_mean_us"] = float((st[:, 7] - st[:, 0]).mean()) / 1e3
    except Exception as e:  # noqa: BLE001
        nms_stages = {"error": repr(e)}
    t = torch.tensor([ms], dtype=torch.float64, device=devc)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * K / (ms_max * 1e-3)
    kern_ms = sum(a.elapsed_time(b) for a, b in ev_k) / K
    launches_per_step = None
    # launches inside the timed region only (the post-region filler loop is excluded)
    L.bg_profile_events(None, None)

    # ---- e2e: pinned host inputs -> H2D -> kernels -> D2H of the result rows, every step ---------
    # through the same pipeline: the copy of a batch runs on its slot's stream, so it overlaps the kernels and the
    # result read of the batches before it (the PCIe link is the bound: 548 MB per step)
    stages = [[torch.empty_like(r) for r in raws_d] for _ in range(depth)]
    d2h_box = [0]

    def e2e_fetch(slot):
        r = pipe.result(slot)
        with torch.cuda.stream(pipe.streams[slot]):
            rows = r.pred_boxes.cpu()
        d2h_box[0] = rows.numel() * 4 + pipe.plans[slot].counts.numel() * 4

    def e2e_steps(n):
        for k in range(n):
            slot = pipe.submitted % pipe.depth
            if k >= depth:
                e2e_fetch(slot)          # the batch that used this slot `depth` steps ago
            st = pipe.streams[slot]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                for sbuf, h in zip(stages[slot], raws_h):
                    sbuf.copy_(h, non_blocking=True)
            pipe.submit(stages[slot])
        for k in range(max(0, n - depth), n):
            e2e_fetch((pipe.submitted - n + k) % pipe.depth)
        pipe.join()

    e2e_steps(max(2, depth))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_steps(K)
    t1.record()
    barrier()
    d2h = d2h_box[0]
    te = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=devc)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * B * K / (float(te.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    alg_bytes = plan.input_bytes  # N*(5+C)*4 per image: the raw head output read once
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    launches_per_step = launches // max(K, 1) if launches else 0
    line = {
        "metric": "images/s (decode+NMS)", "value": value, "unit": "img/s", "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(),
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": int(plan.input_bytes), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "decode_filter_kernel<80> (%s tile loads)" % ("plain" if args.variant == 1 else "TMA bulk"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _ncu_traffic(), "peak_source": peak_src, "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": int(alg_bytes),
                     "whole_step_frac": (alg_bytes / (ms_max / K * 1e-3) / 1e9) / peak},
        "detail": {"batches_in_flight": depth, "single_batch_latency_us": single_ms * 1e3,
                   "kept_rows_per_step": kept_rows, "survivors_per_image": survivors,
                   "launches_per_step": launches_per_step, "image_nms_kernel_stages_us": nms_stages},
    }
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = min(B, max(8, cores))
        ips, threads, tcpu, _ = cpu_oracle_run(sample)
        line["cpu_baseline"] = {"value": ips, "unit": "img/s", "cores": threads, "kind": "port",
                                "sample": "%d of %d images, one pass (%.1f s); CPU oracle port of the reference path, "
                                          "one OpenMP thread per image" % (sample, B, tcpu)}
    if world == 1 and not args.no_extra:
        line["extra"] = extras(torch, ops, synth, devc, peak)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extras(torch, ops, synth, devc, peak):
    """Secondary measurements (not the headline): training-side assignment+loss at config 3 and the
    all-candidates-survive stress case."""
    out = {}
    try:
        B, H, W, C, G = 256, 640, 640, 80, 100
        t = synth.targets(B, G, C, 0).to(devc)
        g = torch.Generator(device=devc).manual_seed(1)
        preds = [torch.randn(B, ny, nx, 3, 5 + C, generator=g, device=devc).requires_grad_(True)
                 for ny, nx in synth.fmap_shapes(H, W)]
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        for _ in range(3):
            loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False)
            loss.backward()
        torch.cuda.synchronize()
        K = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            for p in preds:
                p.grad = None
            loss, _ = ops.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_metrics=False)
            loss.backward()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        alg = B * 10.1e6
        out["train_assign_loss_fwd_bwd_c3"] = {"metric": "images/s (target-assign + loss fwd+bwd), configs[2] on one GPU",
                                               "img_per_s": B / (ms * 1e-3), "ms_per_step": ms, "batch": B, "gt_per_img": G,
                                               "hbm_frac_of_measured": alg / (ms * 1e-3) / 1e9 / peak,
                                               "algorithmic_bytes_per_image": 10.1e6}
        del preds
    except Exception as e:  # noqa: BLE001
        out["train_error"] = repr(e)
    try:
        B, H, W, C = 64, 640, 640, 80
        raws = [r.to(devc) for r in synth.raw_head_outputs(B, H, W, C, "R", 7)]
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (H, W), C, devc, None, 0.65, 0.001, 4)
        plan.enqueue(raws)
        r = plan.result()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.enqueue(raws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["stress_c2R_all_candidates_survive"] = {"img_per_s": B / (ms * 1e-3), "ms_per_step": ms,
                                                    "kept_rows": int(r.pred_boxes.shape[0]),
                                                    "nms_engine": "general segmented engine, grid-pruned pair tests + edge list "
                                                                  "(25,200 survivors per image exceed the per-image kernels)"}
    except Exception as e:  # noqa: BLE001
        out["stress_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="decode kernel: 0 auto, 1 plain loads, 2 TMA bulk")
    ap.add_argument("--depth", type=int, default=4, help="batches in flight (CUDA streams) in the headline loop; 1 = one stream")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extra", action="store_true", help="(default at N=1) also measure the training side and the stress case")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
