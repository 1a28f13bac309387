#!/usr/bin/env python
"""bench.py -- benchmark of the box-geometry hot path on B200.

    python bench.py --gpus 1 --steps 20 --warmup 3             # our arm (CUDA, sm_100a)
    python bench.py --impl reference --steps 2 --warmup 1      # reference arm: the reference's own CPU path

Metric (BASELINE.json): images/s of (i) inference decode + NMS at 640^2, batch 64 (configs[1]) -- the headline
`value` -- and (ii) the training-step target assignment + loss forward/backward (configs[2], batch 256, 100 gt per
image) -- the `train` block -- plus the fraction of the measured HBM peak each sustains.

Headline (decode+NMS).  One "step" = one pass of the fused decode -> score filter -> per-image NMS -> row assembly
over one synthetic batch of 64 images.
  value         device-resident inputs, K steps enqueued through the C ABI, timed with CUDA events;
  e2e           the public API (`ops.DetectPipeline`) with pinned-host inputs: H2D copy of the three head tensors +
                kernels + D2H read of the result rows, every step inside the timed region;
  roofline      the decode+filter kernel timed live with CUDA events recorded around it on its stream;
  cpu_baseline  the reference's own CPU path (unmodified Python + torchvision-CPU from baseline/_ref when present,
                else the C oracle port) on a bounded sample of the same workload;
  parity_check  the CUDA keep-lists of the sampled images against the oracle's, counted and explained.

Training block (`train`, `roofline_train`, `cpu_baseline_train`, `e2e_train`, `gpu_library_baseline`).  One step =
`ops.detection_loss` forward + backward from the head's logits (the form `train_det.py` gets under
`dropin.install()`), batch 256 sharded by image over the ranks (256/N each, strong scaling), followed by the one
tiny all-reduce that turns the per-shard terms into the big-batch loss (`shard.allreduce_loss_terms`, NCCL).
`train.value` is the split form (the head's three conv outputs read in place) replayed from a CUDA graph
(`ops.LossStepGraph`, one launch per step); `train.forms` lists every input form (split / raw / decoded), each
enqueued eagerly from Python (enqueue-bound for small shards) and replayed from a graph.

N > 1 (torchrun): one rank per GPU, barrier + synchronize on both sides, max over ranks.
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
TRAIN = dict(name="c3", B=256, H=640, W=640, C=80, G=100, tseed=0, pseed=1)
PCIE_GEN5_X16_GBS = 63.0


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel="decode_filter_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum of a kernel, per launch, from the committed `ncu --set full`
    capture (profiles/traffic.json); None until a capture of the current kernel exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)[kernel]["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,clocks.mem,clocks.max.mem")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            # nvidia-smi takes ~0.1 s to start (fork + NVML initialisation, which contends with this process's CUDA calls
            # for the driver): wait for its first sample so that the start-up does not fall into the first timed region
            t_end = time.time() + 3.0
            while not self.rows and time.time() < t_end:
                time.sleep(0.01)
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
        sm, mx, reasons, mem, mem_mx = [], 0, set(), [], 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
                if len(r) > 8:
                    mem.append(float(r[7]))
                    mem_mx = max(mem_mx, float(r[8]))
            except Exception:
                pass
        sm.sort()
        mem.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "mem_mhz": mem[len(mem) // 2] if mem else None, "mem_mhz_min": mem[0] if mem else None,
                "mem_max_mhz": mem_mx or None}


def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU baselines are meant to use every host core."""
    threads = os.cpu_count() or 1
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    except OSError:
        pass
    try:
        import torch
        torch.set_num_threads(int(threads))
    except Exception:  # noqa: BLE001
        pass
    return int(threads)


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_oracle_detect(sample_images, steps=1, warmup=0, raws=None):
    """The C oracle port of decode + post_process on `sample_images` images of the headline workload (`raws`: the
    images themselves, e.g. the first ones of the batch the GPU arm processed).
    Returns (img/s, threads, seconds per step, oracle result of the last step)."""
    from oracle import oracle as O
    from vision_conglomerate_b200 import synth
    w = WORKLOAD
    threads = _all_host_threads()
    if raws is None:
        raws = synth.raw_head_outputs(sample_images, w["H"], w["W"], w["C"], w["dist"], w["seed"])
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    times, out = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        preds = O.decode_inference(raws, anc, w["H"], w["W"], None)
        out = O.post_process(preds, w["iou"], w["score"], w["allow"], None)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return sample_images / t, threads, t, out


def cpu_reference_detect(sample_images, steps=1, warmup=0):
    """The UNMODIFIED reference on the host cores: DetectionNet._get_scale_pred x3 + reshape/cat
    (modules/detection.py:69-91) and inference_det.post_process_preds lines 57-97 (torchvision-CPU batched_nms over
    all 25,200 candidates per image), drawing / file output stubbed out.  Returns (img/s, threads, s per step, kept)."""
    import torch
    from oracle import ref_harness
    from vision_conglomerate_b200 import synth
    w = WORKLOAD
    threads = _all_host_threads()
    raws = synth.raw_head_outputs(sample_images, w["H"], w["W"], w["C"], w["dist"], w["seed"])
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ref_harness.load().inference_det.device = "cpu"
    times, kept = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            preds = ref_harness.ref_decode_inference(raws, anc, w["H"], w["W"], None, w["C"])
            cap = ref_harness.ref_post_process(preds, w["C"], w["iou"], w["score"], w["allow"], None)
        dt = time.perf_counter() - t0
        kept = int(sum(len(r) for r in cap["per_image"]))
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return sample_images / t, threads, t, kept


def cpu_train_baseline(sample_images):
    """Loss forward + backward on the host cores for `sample_images` images of the training workload: the unmodified
    reference (DetectionNet._get_scale_pred(inference=False) x3 + DetectionLoss.forward + backward, all threads) when
    baseline/_ref is present, else the scalar C oracle port."""
    import torch
    from oracle import ref_harness
    from vision_conglomerate_b200 import synth
    w = TRAIN
    threads = _all_host_threads()
    t = synth.targets(sample_images, w["G"], w["C"], w["tseed"])
    raws = synth.train_preds(sample_images, w["H"], w["W"], w["C"], w["pseed"])
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    if ref_harness.available():
        ns = ref_harness.load()
        net = ns.DecodeOnly(w["C"])
        loss_mod = ns.DetectionLoss(ns.FakeModel(w["C"], synth.ANCHORS), **synth.LOSS_CONFIG)
        best = None
        for _ in range(2):
            ps = [p.clone().requires_grad_(True) for p in raws]
            t0 = time.perf_counter()
            dec = tuple(net._get_scale_pred(p, a, input_shape=(w["H"], w["W"]), inference=False) for p, a in zip(ps, anc))
            loss, _ = loss_mod(dec, t.clone())
            loss.backward()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return {"value": sample_images / best, "unit": "img/s", "cores": threads, "kind": "reference",
                "sample": "%d of %d images, best of 2 (%.2f s); unmodified _get_scale_pred + DetectionLoss.forward + "
                          "backward on torch-CPU, all host threads" % (sample_images, w["B"], best)}
    from oracle import oracle as O
    t0 = time.perf_counter()
    O.detection_loss(raws, t, anc, synth.LOSS_CONFIG, with_grad=True, input_form="raw")
    dt = time.perf_counter() - t0
    return {"value": sample_images / dt, "unit": "img/s", "cores": 1, "kind": "port",
            "sample": "%d of %d images, one pass (%.2f s); scalar C oracle port" % (sample_images, w["B"], dt)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of decode + post-processing on a bounded sample per
    step (it needs seconds per image: torchvision's CPU NMS runs over all 25,200 candidates of an image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_harness
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 8))   # (every step costs ~4 s per image; reported as run)
    if ref_harness.available():
        sample = 1 if steps * 4 > 40 else 2        # ~4 s per image: keep the whole run within a few minutes
        ips, threads, t, kept = cpu_reference_detect(sample, steps=steps, warmup=warmup)
        kind = "reference"
        what = ("%d of %d images per step; unmodified reference Python (baseline/_ref): _get_scale_pred x3 + cat + "
                "post_process_preds lines 57-97 with torchvision-CPU batched_nms over all 25,200 candidates per image "
                "(drawing / file output stubbed)" % (sample, WORKLOAD["B"]))
    else:
        cores = os.cpu_count() or 1
        sample = min(WORKLOAD["B"], max(8, cores))
        ips, threads, t, _ = cpu_oracle_detect(sample, steps=steps, warmup=warmup)
        kind = "port"
        what = ("%d of %d images per step; C oracle port (oracle/boxgeom_oracle.c), one OpenMP thread per image "
                "(baseline/_ref absent)" % (sample, WORKLOAD["B"]))
    line = {
        "impl": "reference", "metric": "images/s (decode+NMS)", "value": ips, "unit": "img/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t, "images_per_step": sample,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(),
        "cpu_baseline": {"value": ips, "unit": "img/s", "cores": threads, "kind": kind, "sample": what},
        "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def _config():
    w = WORKLOAD
    return {"workload": "configs[1] inference decode+NMS: batch %d at %dx%d, %d classes, 25,200 candidates/img, "
                        "conf %.3f, IoU %.2f, box_allowance %d, trained-like logits (dist T, ~1,700 survivors/img)"
                        % (w["B"], w["H"], w["W"], w["C"], w["score"], w["iou"], w["allow"]),
            "per_gpu_batch": w["B"], "parallelism": "image-sharded replicas, no data-path collective",
            "l2": "inputs (548 MB per step) exceed the 126 MB L2 and two distinct input sets alternate; no explicit flush",
            "pipelining": "consecutive batches in flight on separate CUDA streams (ops.DetectPipeline, --depth); "
                          "every step is a full decode+NMS of its batch into its own output buffers; the per-image NMS runs "
                          "as lean CTAs (512 threads, 46 KB, two per image) on a higher-priority stream next to the resident "
                          "decode CTAs of the following batch"}


# ------------------------------------------------------------------------------------------ our arm
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
    # every rank owns its own image shard (different seed -> different images); two distinct input sets alternate
    raws_h = [r.pin_memory() for r in synth.raw_head_outputs(B, H, W, C, w["dist"], w["seed"] + rank)]
    sets_d = [[r.to(devc) for r in raws_h],
              [r.to(devc) for r in synth.raw_head_outputs(B, H, W, C, w["dist"], w["seed"] + 100 + rank)]]
    raws_d = sets_d[0]
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    plan = ops.DetectPlan([tuple(r.shape) for r in raws_d], anc, (H, W), C, devc, None, w["iou"], w["score"],
                          w["allow"], None, "image", args.variant)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    depth = max(1, args.depth)
    pipe = ops.DetectPipeline([tuple(r.shape) for r in raws_d], anc, (H, W), C, devc, None, w["iou"], w["score"],
                              w["allow"], None, "image", args.variant, depth=depth)

    # warm-up (also settles the workspace / mask budget), every slot of the pipeline and the single-stream plan
    for it in range(max(args.warmup, 3)):
        plan.enqueue(sets_d[it & 1])
        plan.result()
        for _d in range(depth):
            pipe.submit(sets_d[it & 1])
        for d in range(depth):
            pipe.result(d)
    pipe.join()
    plan.enqueue(raws_d)
    det = plan.result()
    kept_rows = int(det.pred_boxes.shape[0])
    survivors = float(det.candidates.float().mean())
    keep_set0 = det.keep_idxs.clone()
    for _d in range(depth):
        pipe.submit(raws_d)
    det_p = pipe.result(depth - 1)
    pipe.join()
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
        # the warm-up steps run again right here: starting the sampler left the GPU idle for ~0.1 s, and the first launches
        # after an idle period are slow (measured: 97-100 instead of 93-94 us per step over 20 steps)
        for i in range(max(args.warmup, 3)):
            pipe.submit(sets_d[i & 1])
        pipe.join()
        barrier()
        launches0 = _lib.launch_count()
        e0.record()
        for i in range(K):
            pipe.submit(sets_d[i & 1])
        pipe.join()
        e1.record()
        launches = _lib.launch_count() - launches0  # kernels of ours enqueued inside the timed region
        barrier()
        # the same K steps on one stream, one batch at a time: the latency view of a step
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for i in range(K):
            plan.enqueue(sets_d[i & 1])
        l1.record()
        barrier()
        single_ms = l0.elapsed_time(l1) / K
        # second pass, untimed as a whole: the same K steps with CUDA events recorded immediately around the
        # decode+filter kernel on its stream (the roofline numerator's duration)
        for i in range(K):
            L.bg_profile_events(ev_k[i][0].cuda_event, ev_k[i][1].cuda_event)
            plan.enqueue(sets_d[i & 1])
        barrier()
        L.bg_profile_events(None, None)
        # ---- the training block runs inside the clock-sampled region as well
        train = train_leg(args, torch, dist, ops, synth, L, devc, world, rank, barrier)
        # the timed regions last milliseconds; keep a load running ~0.5 s more (untimed) so the 100 ms nvidia-smi
        # sampler sees clocks under this load
        t_end = time.time() + 0.5
        while time.time() < t_end:
            plan.enqueue(raws_d)
            torch.cuda.synchronize()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # stage breakdown of the per-image NMS kernels (one extra untimed step each with the %globaltimer hook armed): the
    # 1024-thread kernel of the single-batch plan and the lean kernel of the pipelined plans, each running alone
    def nms_stage_times(pl):
        try:
            ns = int(L.bg_profile_stamps_per_image())
            stamps = torch.zeros(B, ns, dtype=torch.int64, device=devc)
            L.bg_profile_stamps(stamps.data_ptr())
            pl.enqueue(raws_d)
            torch.cuda.synchronize()
            L.bg_profile_stamps(None)
            st = stamps.cpu().double()
            if float(st[:, 7].min()) <= 0:
                return None
            names = ["load_slots", "grid_bucket", "pair_tests", "resolve", "sort", "rank+lookback", "write_rows"]
            out = {n: float((st[:, i + 1] - st[:, i]).mean()) / 1e3 for i, n in enumerate(names)}
            out["kernel_span_us"] = float(st[:, 7].max() - st[:, 0].min()) / 1e3
            out["per_image_mean_us"] = float((st[:, 7] - st[:, 0]).mean()) / 1e3
            return out
        except Exception as e:  # noqa: BLE001
            return {"error": repr(e)}

    nms_stages = nms_stage_times(plan)
    nms_stages_lean = None
    if pipe.plans[0].params.nms_path == 5:
        lean_plan = ops.DetectPlan([tuple(r.shape) for r in raws_d], anc, (H, W), C, devc, None, w["iou"], w["score"],
                                   w["allow"], None, "image", args.variant, "per_image_lean")
        lean_plan.enqueue(raws_d)
        lean_plan.result()
        nms_stages_lean = nms_stage_times(lean_plan)
    t = torch.tensor([ms], dtype=torch.float64, device=devc)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * K / (ms_max * 1e-3)
    kern_ms = sum(a.elapsed_time(b) for a, b in ev_k) / K

    # ---- e2e: pinned host inputs -> H2D -> kernels -> D2H of the result rows, every step ---------
    # through the same pipeline: the copy of a batch runs on its slot's stream, so it overlaps the kernels and the
    # result read of the batches before it (the host link is the bound: 548 MB per step)
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
    e2e_ms = float(te.item())
    e2e = world * B * K / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    alg_bytes = plan.input_bytes  # N*(5+C)*4 per image: the raw head output read once
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    launches_per_step = launches // max(K, 1) if launches else 0
    h2d_gbs = world * plan.input_bytes * K / (e2e_ms * 1e-3) / 1e9
    line = {
        "metric": "images/s (decode+NMS)", "value": value, "unit": "img/s", "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(),
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": int(plan.input_bytes), "d2h_bytes_per_step": int(d2h),
                "bound": "host link: %.1f GB/s of H2D aggregate over %d GPU(s) = %.1f GB/s per GPU (PCIe Gen5 x16 ~%.0f GB/s "
                         "per GPU, shared host memory bandwidth across GPUs); in inference_det.py the head tensors are "
                         "produced on the device, so this is the worst case" % (h2d_gbs, world, h2d_gbs / world, PCIE_GEN5_X16_GBS),
                "h2d_gbs_aggregate": h2d_gbs},
        "gpu_launches": int(launches + train.get("launches_in_timed_region", 0)),
        "roofline": {"bound": "hbm", "kernel": "decode_filter_kernel<80> (%s tile loads)" % ("plain" if args.variant == 1 else "TMA bulk"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _ncu_traffic(), "peak_source": peak_src, "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": int(alg_bytes),
                     "whole_step_frac": (alg_bytes / (ms_max / K * 1e-3) / 1e9) / peak},
        "detail": {"batches_in_flight": depth, "single_batch_latency_us": single_ms * 1e3,
                   "kept_rows_per_step": kept_rows, "survivors_per_image": survivors,
                   "launches_per_step": launches_per_step, "image_nms_kernel_stages_us": nms_stages,
                   "pipelined_nms": {"nms_path": int(pipe.plans[0].params.nms_path), "kernel": "image_nms_kernel<InmsLean> (512 threads, "
                                     "46 KB, main + helper CTA per image) on a higher-priority stream"
                                     if pipe.plans[0].params.nms_path == 5 else "image_nms_kernel (1024 threads, one CTA per image)",
                                     "stages_us_running_alone": nms_stages_lean}},
    }
    for k in ("train", "roofline_train", "e2e_train"):
        if k in train:
            line[k] = train[k]
    if world == 1 and not args.no_cpu:
        line.update(cpu_legs(torch, keep_set0, B, raws_h))
    if world == 1 and not args.no_extra:
        line["gpu_library_baseline"] = gpu_library_baseline(torch, ops, synth, devc)
        line["extra"] = extras(torch, ops, synth, devc)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def train_leg(args, torch, dist, ops, synth, L, devc, world, rank, barrier):
    """configs[2]: batch 256 at 640^2, 100 gt/img, image-sharded over the ranks; fwd + bwd + normaliser all-reduce."""
    from vision_conglomerate_b200 import _lib, shard
    out = {}
    w = TRAIN
    Bf, H, W, C, G = w["B"], w["H"], w["W"], w["C"], w["G"]
    s, e = shard.shard_range(Bf, world, rank)
    Bl = e - s
    anc = [synth.anchors_tensor(sc) for sc in synth.SCALES]
    cfg = dict(synth.LOSS_CONFIG, num_classes=C)
    t_full = synth.targets(Bf, G, C, w["tseed"]).to(devc)
    g = torch.Generator(device=devc).manual_seed(w["pseed"])
    full = [torch.randn(Bf, ny, nx, 3, 5 + C, generator=g, device=devc) for ny, nx in synth.fmap_shapes(H, W)]
    # the single-GPU big-batch loss every sharded run must reproduce (untimed)
    with torch.no_grad():
        loss_full, _ = ops.detection_loss(full, t_full, anc, cfg, with_metrics=False, input_form="raw")
    loss_full = float(loss_full)
    t_loc = shard.shard_targets(t_full, s, e)
    logits = [x[s:e].clone().requires_grad_(True) for x in full]
    del full
    torch.cuda.empty_cache()
    cells = [Bl * ny * nx * 3 for ny, nx in synth.fmap_shapes(H, W)]   # this shard's cells per scale (summed by the all-reduce)
    K = args.steps

    def step(form_inputs, form):
        for p in form_inputs:
            for q in (p if isinstance(p, tuple) else (p,)):
                q.grad = None
        loss, _, sc = ops.detection_loss(form_inputs, t_loc, anc, cfg, with_metrics=False, return_scalars=True, input_form=form)
        loss.backward()
        # per-shard terms -> big-batch loss: one 15-double all-reduce over NCCL (no-op at world 1)
        return shard.allreduce_loss_terms(sc, cells, cfg)

    samples = {}

    def timed(form_inputs, form, steps, regions=3):
        """`regions` timed regions of exactly `steps` steps each (barrier + synchronize on both sides, max over ranks);
        the best region is reported, all of them are listed (`region_ms_per_step`): the eager loop is enqueue-bound
        for small shards and a busy host shows up as an occasional slow region."""
        for _ in range(max(args.warmup, 3)):
            comb = step(form_inputs, form)
        per = []
        for _r in range(regions):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = _lib.launch_count()
            a.record()
            for _ in range(steps):
                comb = step(form_inputs, form)
            b.record()
            nl = _lib.launch_count() - l0
            barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=devc)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            per.append(float(t.item()) / steps)
        samples[form] = per
        return min(per), comb, nl

    def graphed(form_inputs, form):
        """The same step captured in a CUDA graph (ops.LossStepGraph: fwd + bwd + pack + NCCL all-reduce + combine in one
        launch): small shards are enqueue-bound when run eagerly, the replay is bound by the kernels and the collective."""
        clones = [tuple(q.detach().clone().requires_grad_(True) for q in p) if isinstance(p, tuple)
                  else p.detach().clone().requires_grad_(True) for p in form_inputs]
        gs = ops.LossStepGraph(clones, t_loc, anc, cfg, input_form=form, cells=cells)
        for _ in range(3):
            gs.replay()
        per = []
        for _r in range(3):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(K):
                gs.replay()
            b.record()
            barrier()
            tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=devc)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            per.append(float(tt.item()) / K)
        g_rel = abs(float(gs.combined) - loss_full) / abs(loss_full)
        if g_rel > 1e-6:
            raise RuntimeError("graphed sharded loss (%s) differs from the big-batch loss (rel %.2e)" % (form, g_rel))
        return min(per), per, float(gs.combined), g_rel

    def fill_kernel_ms(form_inputs, form):
        """The dense-gradient fill alone (events around it inside bg_loss_bwd, bg_profile_events_loss)."""
        leaves = [q for p in form_inputs for q in (p if isinstance(p, tuple) else (p,))]
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(K, 10))]
        for x, y in evs:
            x.record()
            y.record()
        for x, y in evs:
            for p in leaves:
                p.grad = None
            loss, _ = ops.detection_loss(form_inputs, t_loc, anc, cfg, with_metrics=False, input_form=form)
            L.bg_profile_events_loss(x.cuda_event, y.cuda_event)
            loss.backward()
        barrier()
        L.bg_profile_events_loss(None, None)
        return sum(x.elapsed_time(y) for x, y in evs) / len(evs)

    # counts for the algorithmic bytes
    with torch.no_grad():
        _, _, sc = ops.detection_loss([x.detach() for x in logits], t_loc, anc, cfg, with_metrics=False, return_scalars=True,
                                      input_form="raw")
    M_loc = float(sc[:, 6].sum())
    N = synth.candidates_per_image(H, W)
    D = 5 + C
    # SURVEY 8d, fwd+bwd per shard: objectness plane + matched rows + targets (fwd); dense gradient + residual + matched rows (bwd)
    alg = Bl * N * 4 + M_loc * D * 4 + t_loc.shape[0] * 24 + Bl * N * D * 4 + Bl * N * 4 + M_loc * D * 4
    fill_alg = Bl * N * D * 4 + Bl * N * 4
    peak, peak_src = _peaks()

    # the three input forms on the same shard: eager (3 regions of K steps) and graph replay (3 regions of K replays)
    what = {"split": "SURVEY 8 f3: the head's three conv outputs (conf / cls / bbox) read in place -- what train_det.py gets under "
                     "dropin.install(EffiDecHead=...) (zero-copy with a channels-last model)",
            "raw": "the head's interleaved rows (logits), training-mode decode fused -- dropin.install() without the head patch",
            "decoded": "the reference's contract: tensors already decoded by _get_scale_pred"}
    inputs = {"raw": logits,
              "split": [tuple(y.contiguous().requires_grad_(True) for y in (x.detach()[..., 0], x.detach()[..., 1:1 + C], x.detach()[..., 1 + C:]))
                        for x in logits],
              "decoded": [ops.decode_scale(x.detach(), a_, (H, W), False).requires_grad_(True) for x, a_ in zip(logits, anc)]}
    forms, nl_by = {}, {}
    for form in ("split", "raw", "decoded"):
        ent = {"what": what[form]}
        try:
            ms_e, comb, nl = timed(inputs[form], form, K)
            nl_by[form] = nl
            rel = abs(float(comb) - loss_full) / abs(loss_full) if form != "decoded" else None
            if rel is not None and rel > 1e-6:
                raise RuntimeError("sharded loss (%s) %.9f differs from the big-batch loss %.9f (rel %.2e)" % (form, float(comb), loss_full, rel))
            ent["eager"] = {"ms_per_step": ms_e, "img_per_s": Bf / (ms_e * 1e-3), "hbm_frac": alg / (ms_e * 1e-3) / 1e9 / peak,
                            "region_ms_per_step": samples[form], "launches_per_step": nl // max(K, 1),
                            "sharded_allreduced_loss": float(comb), "loss_rel_err_vs_big_batch": rel}
        except Exception as ex:  # noqa: BLE001
            ent["eager"] = {"error": repr(ex)}
        try:
            ms_g, per_g, comb_g, rel_g = graphed(inputs[form], form) if form != "decoded" else (None, None, None, None)
            if ms_g is not None:
                ent["graph"] = {"ms_per_step": ms_g, "img_per_s": Bf / (ms_g * 1e-3), "hbm_frac": alg / (ms_g * 1e-3) / 1e9 / peak,
                                "region_ms_per_step": per_g, "sharded_allreduced_loss": comb_g, "loss_rel_err_vs_big_batch": rel_g}
        except Exception as ex:  # noqa: BLE001
            ent["graph"] = {"error": repr(ex)}
        forms[form] = ent
    primary = "split"
    best = forms[primary].get("graph") if "ms_per_step" in forms[primary].get("graph", {}) else forms[primary]["eager"]
    mode = "cuda_graph" if best is forms[primary].get("graph") else "eager"
    ms_best = best["ms_per_step"]
    # forward alone (eager, primary form)
    detached = [tuple(q.detach() for q in p) for p in inputs[primary]]
    for _ in range(3):
        ops.detection_loss(detached, t_loc, anc, cfg, with_metrics=False, input_form=primary)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        ops.detection_loss(detached, t_loc, anc, cfg, with_metrics=False, input_form=primary)
    b.record()
    barrier()
    fwd_ms = a.elapsed_time(b) / K
    keep_mode = ops.PRECLEAR_SPLIT_GRADS
    ops.PRECLEAR_SPLIT_GRADS = False      # (for this measurement the clear runs inside the backward, between the events)
    fill_split = fill_kernel_ms(inputs["split"], "split")
    ops.PRECLEAR_SPLIT_GRADS = keep_mode
    fill_raw = fill_kernel_ms(inputs["raw"], "raw")
    out["train"] = {
        "metric": "images/s (target-assign + loss fwd+bwd)", "value": Bf / (ms_best * 1e-3), "unit": "img/s",
        "ms_per_step": ms_best, "n_gpus": world, "scaling": "strong", "global_batch": Bf, "per_gpu_batch": Bl,
        "gt_per_img": G, "input_form": primary + ": " + what[primary],
        "mode": ("cuda_graph: ops.LossStepGraph replays fwd + bwd + loss-term pack + all-reduce + combine captured once (fixed "
                 "shapes and addresses); the eager numbers of every form are in train.forms") if mode == "cuda_graph" else "eager",
        "timing": "best of 3 timed regions of %d steps each (all listed in region_ms_per_step)" % K,
        "region_ms_per_step": best["region_ms_per_step"], "forward_ms_eager": fwd_ms,
        "launches_per_step_eager": nl_by.get(primary, 0) // max(K, 1),
        "collective": "one all-reduce(SUM) of 15 float64 per step (shard.allreduce_loss_terms, %s)" % ("NCCL" if world > 1 else "world 1: skipped"),
        "loss_check": {"big_batch_loss": loss_full, "sharded_allreduced_loss": best["sharded_allreduced_loss"],
                       "rel_err": best["loss_rel_err_vs_big_batch"], "tol": 1e-6},
        "workload": "configs[2]: batch %d at %dx%d, %d gt/img, CIoU + objectness/class BCE, image-sharded %d per GPU"
                    % (Bf, H, W, G, Bl),
        "forms": forms,
    }
    # kernels of ours inside the timed regions of this block: the last eager region of every form (host-side counter) plus
    # the graph replays (3 regions x K replays x the launches one step holds; replays do not pass through the host counter)
    replayed = sum(3 * K * (nl_by.get(f, 0) // max(K, 1)) for f in ("split", "raw") if "ms_per_step" in forms[f].get("graph", {}))
    out["launches_in_timed_region"] = int(sum(nl_by.values()) + replayed)
    out["roofline_train"] = {
        "bound": "hbm", "kernel": "whole step (fwd+bwd, %s, %s form); dominant part: the dense-gradient fill" % (mode, primary),
        "achieved": alg / (ms_best * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms_best * 1e-3) / 1e9 / peak,
        "algorithmic_bytes_per_step": int(alg), "algorithmic_bytes_per_image": alg / Bl, "peak_source": peak_src,
        "traffic": _ncu_traffic("train_step_split"),
        "dominant_kernel": {"name": "cudaMemsetAsync of the class / box gradient planes + loss_bwd_conf_kernel (split form; timed "
                                    "alone here -- in the step the clear runs on the caller's stream next to the forward kernels, "
                                    "which run on a higher-priority stream)",
                            "kernel_ms": fill_split, "algorithmic_bytes_per_launch": int(fill_alg),
                            "achieved": fill_alg / (fill_split * 1e-3) / 1e9, "frac": fill_alg / (fill_split * 1e-3) / 1e9 / peak},
        "interleaved_fill_kernel": {"name": "l2_pin_kernel + loss_bwd_stream_kernel (raw / decoded forms)", "kernel_ms": fill_raw,
                                    "algorithmic_bytes_per_launch": int(fill_alg), "achieved": fill_alg / (fill_raw * 1e-3) / 1e9,
                                    "frac": fill_alg / (fill_raw * 1e-3) / 1e9 / peak, "traffic": _ncu_traffic("loss_bwd_stream_kernel")},
        "survey_10.1MB_per_image_frac": Bl * 10.1e6 / (ms_best * 1e-3) / 1e9 / peak,
        "interleaved_forms_frac": {f: forms[f].get("graph", forms[f]["eager"]).get("hbm_frac") for f in ("raw", "decoded")},
    }
    del inputs["decoded"]
    # ---- weak scaling of the same step (N > 1): every rank runs a FULL batch of its own (per-GPU work fixed as N grows),
    #      same captured step incl. the NCCL all-reduce of the loss terms; the strong-scaling numbers above are configs[2]
    if world > 1:
        try:
            gw = torch.Generator(device=devc).manual_seed(w["pseed"] + 1000 + rank)
            t_w = synth.targets(Bf, G, C, w["tseed"] + rank).to(devc)
            split_w = []
            for ny, nx in synth.fmap_shapes(H, W):
                split_w.append(tuple(torch.randn(Bf, ny, nx, 3, n_, generator=gw, device=devc).squeeze(-1).contiguous().requires_grad_(True)
                                     if n_ == 1 else torch.randn(Bf, ny, nx, 3, n_, generator=gw, device=devc).requires_grad_(True)
                                     for n_ in (1, C, 4)))
            cells_w = [Bf * ny * nx * 3 for ny, nx in synth.fmap_shapes(H, W)]
            gsw = ops.LossStepGraph(split_w, t_w, anc, cfg, input_form="split", cells=cells_w)
            for _ in range(3):
                gsw.replay()
            per = []
            for _r in range(3):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(K):
                    gsw.replay()
                b.record()
                barrier()
                tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=devc)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                per.append(float(tt.item()) / K)
            ms_w = min(per)
            lw = float(gsw.combined)
            if not (lw == lw and abs(lw) < 1e6):
                raise RuntimeError("weak-scaling step: combined loss %r" % lw)
            out["train"]["weak_scaling"] = {
                "value": world * Bf / (ms_w * 1e-3), "unit": "img/s", "ms_per_step": ms_w, "per_gpu_batch": Bf, "global_batch": world * Bf,
                "scaling": "weak", "region_ms_per_step": per, "combined_loss_of_the_global_batch": lw,
                "what": "every rank runs its own batch of %d images (split form, CUDA-graph replay incl. the NCCL all-reduce of the "
                        "15 loss terms); max over ranks" % Bf}
            del gsw, split_w
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            out["train"]["weak_scaling"] = {"error": repr(ex)}
    # ---- e2e_train: pinned host conv outputs -> H2D -> fwd+bwd -> D2H of the loss, every step (primary form)
    try:
        host = [tuple(q.detach().cpu().pin_memory() for q in p) for p in inputs[primary]]
        th = t_loc.cpu().pin_memory()
        stage = [tuple(torch.empty_like(q).requires_grad_(True) for q in p) for p in inputs[primary]]
        tdev = torch.empty_like(t_loc)
        Ke = max(2, min(K, 5))

        def e2e_step():
            with torch.no_grad():
                for sp, hp in zip(stage, host):
                    for sbuf, h in zip(sp, hp):
                        sbuf.copy_(h, non_blocking=True)
                tdev.copy_(th, non_blocking=True)
            for sp in stage:
                for q in sp:
                    q.grad = None
            loss, _ = ops.detection_loss(stage, tdev, anc, cfg, with_metrics=False, input_form=primary)
            loss.backward()
            return float(loss.detach())         # D2H of the step's result

        e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(Ke):
            e2e_step()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=devc)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e = float(t.item()) / Ke
        h2d = sum(q.numel() * 4 for p in host for q in p) + th.numel() * 4
        out["e2e_train"] = {"value": Bf / (ms_e * 1e-3), "unit": "img/s", "ms_per_step": ms_e, "h2d_bytes_per_step": int(h2d),
                            "d2h_bytes_per_step": 4, "steps": Ke,
                            "bound": "host link (%.1f GB/s H2D per GPU); under train_det.py the head's outputs are produced on the "
                                     "device, so this is the worst case" % (h2d / (ms_e * 1e-3) / 1e9)}
    except Exception as ex:  # noqa: BLE001
        out["e2e_train"] = {"error": repr(ex)}
    return out


def cpu_legs(torch, keep_gpu, B, raws_h):
    """cpu_baseline (+ parity_check against the oracle on the same images) and cpu_baseline_train."""
    from oracle import ref_harness
    from oracle.parity import explain_keep_mismatches
    from vision_conglomerate_b200 import synth
    out = {}
    cores = os.cpu_count() or 1
    sample = min(B, max(8, cores))
    ips, threads, tcpu, ref = cpu_oracle_detect(sample, raws=[r[:sample].contiguous() for r in raws_h])
    N = synth.candidates_per_image(WORKLOAD["H"], WORKLOAD["W"])
    kg = keep_gpu.cpu().numpy()
    kg = kg[kg < sample * N]
    par = explain_keep_mismatches(ref["score"], ref["xyxy"], N, ref["keep"], kg, WORKLOAD["iou"], WORKLOAD["score"])
    out["parity_check"] = {"images": sample, "kept": par["kept_ref"], "kept_cuda": par["kept_gpu"], "mismatches": par["mismatches"],
                           "explained": {k: par[k] for k in ("score_threshold", "iou_at_threshold", "score_tie", "cascade")},
                           "unexplained": len(par["unexplained"]),
                           "checker": "C oracle port on the first %d images of the headline batch (oracle/parity.py)" % sample}
    port = {"value": ips, "unit": "img/s", "cores": threads, "kind": "port",
            "sample": "%d of %d images, one pass (%.1f s); CPU oracle port of the reference path, one OpenMP thread per image"
                      % (sample, B, tcpu)}
    if ref_harness.available():
        n_ref = 2
        rips, rthreads, rt, rkept = cpu_reference_detect(n_ref)
        out["cpu_baseline"] = {"value": rips, "unit": "img/s", "cores": rthreads, "kind": "reference",
                               "sample": "%d of %d images, one pass (%.1f s); unmodified reference Python + torchvision-CPU "
                                         "batched_nms over all 25,200 candidates per image" % (n_ref, B, rt),
                               "port": port}
    else:
        out["cpu_baseline"] = port
    try:
        out["cpu_baseline_train"] = cpu_train_baseline(8)
    except Exception as ex:  # noqa: BLE001
        out["cpu_baseline_train"] = {"error": repr(ex)}
    return out


def gpu_library_baseline(torch, ops, synth, devc):
    """The 'existing sm_100 kernels' bar: the reference's own torch-CUDA loss and torchvision's CUDA
    batched_nms on this B200, through the unmodified reference code (baseline/_ref)."""
    from oracle import ref_harness
    out = {}
    if not ref_harness.available():
        return {"unavailable": "baseline/_ref is not present on this box"}
    try:
        import torchvision
        ns = ref_harness.load()
        w = TRAIN
        Bt = 64   # the reference's loss makes ~30 host syncs per call: its time barely depends on the batch
        t = synth.targets(Bt, w["G"], w["C"], w["tseed"]).to(devc)
        g = torch.Generator(device=devc).manual_seed(w["pseed"])
        raws = [torch.randn(Bt, ny, nx, 3, 5 + w["C"], generator=g, device=devc).requires_grad_(True)
                for ny, nx in synth.fmap_shapes(w["H"], w["W"])]
        net = ns.DecodeOnly(w["C"])
        model = ns.FakeModel(w["C"], synth.ANCHORS).to(devc)
        loss_mod = ns.DetectionLoss(model, **synth.LOSS_CONFIG)
        anc = [model.sm_anchors.data, model.md_anchors.data, model.lg_anchors.data]

        def ref_step():
            for p in raws:
                p.grad = None
            dec = tuple(net._get_scale_pred(p, a, input_shape=(w["H"], w["W"]), inference=False) for p, a in zip(raws, anc))
            loss, _ = loss_mod(dec, t)
            loss.backward()

        ref_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ref_step()
        torch.cuda.synchronize()
        ref_ms = (time.perf_counter() - t0) / 3 * 1e3

        def our_step():
            for p in raws:
                p.grad = None
            loss, _ = ops.detection_loss(raws, t, anc, synth.LOSS_CONFIG, input_form="raw")   # with the metrics read, like the reference
            loss.backward()

        our_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            our_step()
        torch.cuda.synchronize()
        our_ms = (time.perf_counter() - t0) / 10 * 1e3
        out["loss_fwd_bwd"] = {"batch": Bt, "reference_torch_cuda_ms": ref_ms, "ours_ms": our_ms, "speedup": ref_ms / our_ms,
                               "what": "unmodified _get_scale_pred(inference=False) x3 + DetectionLoss.forward + backward on "
                                       "torch-CUDA vs ops.detection_loss(raw) + backward, both including the metrics dict"}
        del raws
    except Exception as ex:  # noqa: BLE001
        out["loss_error"] = repr(ex)
    try:
        import torchvision
        w = WORKLOAD
        B = 8
        raws = [r.to(devc) for r in synth.raw_head_outputs(B, w["H"], w["W"], w["C"], w["dist"], w["seed"])]
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        ns = ref_harness.load()
        ns.inference_det.device = "cuda"
        with torch.no_grad():
            preds = ref_harness.ref_decode_inference(raws, [a.to(devc) for a in anc], w["H"], w["W"], None, w["C"])
            ref_harness.ref_post_process(preds, w["C"], w["iou"], w["score"], w["allow"], None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            preds = ref_harness.ref_decode_inference(raws, [a.to(devc) for a in anc], w["H"], w["W"], None, w["C"])
            cap = ref_harness.ref_post_process(preds, w["C"], w["iou"], w["score"], w["allow"], None)
            torch.cuda.synchronize()
            ref_ms = (time.perf_counter() - t0) * 1e3
        plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (w["H"], w["W"]), w["C"], devc, None, w["iou"], w["score"], w["allow"])
        plan.enqueue(raws)
        plan.result()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            plan.enqueue(raws)
            r = plan.result()
        torch.cuda.synchronize()
        our_ms = (time.perf_counter() - t0) / 10 * 1e3
        out["decode_nms"] = {"batch": B, "reference_torch_cuda_ms": ref_ms, "ours_ms": our_ms, "speedup": ref_ms / our_ms,
                             "reference_img_per_s": B / (ref_ms * 1e-3), "ours_img_per_s": B / (our_ms * 1e-3),
                             "what": "unmodified decode + post_process_preds (ATen + torchvision-CUDA batched_nms over all "
                                     "candidates, incl. its per-image host loop with drawing stubbed) vs ops.detect, "
                                     "wall clock incl. the result read; batch %d of the headline workload" % B,
                             "kept_rows": [int(sum(len(x) for x in cap["per_image"])), int(r.pred_boxes.shape[0])]}
    except Exception as ex:  # noqa: BLE001
        out["decode_nms_error"] = repr(ex)
    return out


def extras(torch, ops, synth, devc):
    """Secondary measurement: the all-candidates-survive stress case (dist R)."""
    out = {}
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
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    # ---- BASELINE configs[4]: batch-1 video frames at 640^2, IoU 0.35 / score 0.3, tracked classes, og (720, 1280); wall clock
    try:
        frames = [[r.to(devc) for r in synth.raw_head_outputs(1, 640, 640, 80, "TP", 7 + f)] for f in range(16)]
        shapes = [tuple(r.shape) for r in frames[0]]
        tr = synth.tracked_classes_default()
        res = {}
        for name, host in (("device_rows_count_read_and_sync", False), ("host_rows_polled_flag", True)):
            plan = ops.DetectPlan(shapes, anc, (640, 640), 80, devc, (720, 1280), 0.35, 0.3, 4, tr, host_result=host)
            lat, rows = [], 0
            for f in range(332):
                t0 = time.perf_counter()
                plan.enqueue(frames[f % 16])
                rows = plan.result_host().rows.shape[0] if host else plan.result().pred_boxes.shape[0]
                lat.append((time.perf_counter() - t0) * 1e6)
            lat = sorted(lat[32:])
            res[name] = {"p50_us": lat[len(lat) // 2], "p99_us": lat[int(len(lat) * 0.99)], "frames": len(lat), "rows_last_frame": int(rows)}
        res["what"] = ("per frame: enqueue the two kernels + get the result, wall clock; host_rows_polled_flag = DetectPlan(host_result=True): "
                       "the kernels write rows / counts into page-locked host memory and store a flag last (no copy, no stream sync)")
        out["c5_batch1_latency"] = res
    except Exception as e:  # noqa: BLE001
        out["c5_error"] = repr(e)
    # ---- BASELINE configs[3]: 1280^2 (100,800 candidates/img), batch 32: decode+NMS (trained-like logits drawn on the device) and
    #      assignment + loss fwd+bwd with 300 gt/img
    try:
        B, S, C = 32, 1280, 80
        g = torch.Generator(device=devc).manual_seed(9)
        raws = []
        for ny, nx in synth.fmap_shapes(S, S):
            z = torch.randn(B, ny, nx, 3, 5 + C, generator=g, device=devc)
            z[..., 0] = -9.0 + 2.0 * z[..., 0]
            z[..., 1:1 + C] = -4.0 + 1.5 * z[..., 1:1 + C]
            raws.append(z.contiguous())
        shapes = [tuple(r.shape) for r in raws]
        plan = ops.DetectPlan(shapes, anc, (S, S), C, devc, None, 0.65, 0.001, 4)
        plan.enqueue(raws)
        r = plan.result()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            plan.enqueue(raws)
        e1.record()
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1) / 5
        pipe = ops.DetectPipeline(shapes, anc, (S, S), C, devc, None, 0.65, 0.001, 4, None, depth=4)
        for _ in range(2):
            for _d in range(4):
                pipe.submit(raws)
            for d in range(4):
                pipe.result(d)
        pipe.join()
        e0.record()
        for _ in range(16):
            pipe.submit(raws)
        pipe.join()
        e1.record()
        torch.cuda.synchronize()
        msp = e0.elapsed_time(e1) / 16
        peak, _src = _peaks()
        out["c4_detect_1280_b32"] = {"ms_per_batch": ms1, "img_per_s": B / (ms1 * 1e-3), "ms_per_batch_4_in_flight": msp,
                                     "img_per_s_4_in_flight": B / (msp * 1e-3), "hbm_frac_4_in_flight": plan.input_bytes / (msp * 1e-3) / 1e9 / peak,
                                     "survivors_per_image": float(r.candidates.float().mean()), "kept_rows": int(r.pred_boxes.shape[0])}
        del raws, plan, pipe
        torch.cuda.empty_cache()
        t = synth.targets(B, 300, C, 0).to(devc)
        parts = [tuple(torch.randn(B, ny, nx, 3, n_, generator=g, device=devc).squeeze(-1).contiguous().requires_grad_(True) if n_ == 1
                       else torch.randn(B, ny, nx, 3, n_, generator=g, device=devc).requires_grad_(True) for n_ in (1, C, 4))
                 for ny, nx in synth.fmap_shapes(S, S)]
        cfg = dict(synth.LOSS_CONFIG, num_classes=C)
        gs = ops.LossStepGraph(parts, t, anc, cfg, input_form="split")
        for _ in range(3):
            gs.replay()
        e0.record()
        for _ in range(10):
            gs.replay()
        e1.record()
        torch.cuda.synchronize()
        mst = e0.elapsed_time(e1) / 10
        N4 = synth.candidates_per_image(S, S)
        out["c4_train_1280_b32_300gt"] = {"ms_per_step": mst, "img_per_s": B / (mst * 1e-3), "form": "split, CUDA-graph replay",
                                          "hbm_frac_dense_gradient_only": B * N4 * (5 + C) * 4 / (mst * 1e-3) / 1e9 / peak}
        del gs, parts
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["c4_error"] = repr(e)
    # ---- SURVEY 8 f2, inference side: boolean masks of 200 kept rows, protos 160^2 -> 640^2 (inference_seg.py:115-117)
    try:
        Bm, K, Hp, Wp, H, W = 8, 32, 160, 160, 640, 640
        g = torch.Generator(device=devc).manual_seed(5)
        counts = torch.full((Bm,), 25)
        coefs = torch.tanh(torch.randn(int(counts.sum()), K, generator=g, device=devc))
        protos = torch.randn(Bm, K, Hp, Wp, generator=g, device=devc)

        def torch_masks():
            r0 = 0
            for i, c in enumerate(counts.tolist()):
                m = (coefs[r0:r0 + c] @ protos[i].reshape(K, -1)).reshape(-1, Hp, Wp).sigmoid()
                m = torch.nn.functional.interpolate(m.unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False)
                torch.gt(m, 0.5)
                r0 += c

        def timed(fn, n=20):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        a, b = timed(lambda: ops.seg_masks(coefs, counts, protos, (H, W))), timed(torch_masks)
        out["seg_masks_200x640x640"] = {"ours_ms": a, "same_torch_calls_ms": b, "speedup": b / a, "output_MB": 200 * H * W / 1e6}
    except Exception as e:  # noqa: BLE001
        out["seg_masks_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="decode kernel: 0 auto, 1 plain loads, 2 TMA bulk")
    ap.add_argument("--depth", type=int, default=4, help="batches in flight (CUDA streams) in the headline loop; 1 = one stream")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity legs")
    ap.add_argument("--no-extra", action="store_true", help="skip gpu_library_baseline and the stress case")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
