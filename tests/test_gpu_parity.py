"""GPU parity suite (-m gpu): the CUDA path, called through the Python shim over the C ABI, against the
CPU oracle on the same seeded inputs and against the committed golden fixtures (reference outputs).

Bars: integer / index outputs bit-exact; fp32 outputs rtol 1e-5 (atol stated per test)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle.parity import explain_keep_mismatches, summarize
from vision_conglomerate_b200 import synth
from tests.util import (ASSIGN_VARIANTS, assign_variant_case, assert_close, canon, digest, golden, rows_canon, rows_order,
                        seg_extra_columns)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vision_conglomerate_b200 import ops as _ops
    return _ops


def dev(t):
    return t.cuda() if isinstance(t, torch.Tensor) else torch.as_tensor(t).cuda()


# ------------------------------------------------------------------------------------------ NMS (B4)
@pytest.mark.parametrize("case", ["rand", "ties", "dense1", "hand_thr05", "hand_thr05m", "hand_thr0"])
def test_nms_golden(ops, case):
    """Identical fp32 inputs -> keep list bit-exact with torchvision-CPU (the reference's NMS)."""
    g = golden("nms")
    b, s, i, thr = g[case + "_boxes"], g[case + "_scores"], g[case + "_idxs"], float(g[case + "_thr"])
    keep = ops.batched_nms(dev(b), dev(s), dev(i), thr).cpu().numpy()
    assert np.array_equal(keep, canon(g[case + "_keep"], s))


@pytest.mark.parametrize("n,groups,thr,ties", [(1, 1, 0.5, False), (2, 2, 0.5, False), (63, 1, 0.3, False),
                                               (64, 1, 0.3, True), (65, 3, 0.7, False), (4097, 2, 0.5, True),
                                               (20000, 8, 0.45, False), (30000, 1, 0.6, False),
                                               (50000, 700, 0.5, True)])
def test_nms_oracle(ops, n, groups, thr, ties):
    b, s, i = synth.nms_boxes(n, groups, seed=n, ties=ties)
    i = i * 3 - 7  # arbitrary (negative, gapped) group ids
    ref = O.batched_nms(b, s, i, thr)
    keep = ops.batched_nms(dev(b), dev(s), dev(i), thr).cpu().numpy()
    assert np.array_equal(keep, ref)


@pytest.mark.parametrize("n,groups,thr", [(7000, 1, 0.5), (9000, 3, 0.3), (5000, 2, 0.04), (5000, 2, 0.0)])
def test_nms_crowded_and_tiny_thresholds(ops, n, groups, thr):
    """Heavily overlapping clusters overflow the overlap-edge list (the dense bit-matrix fallback takes over inside
    the same call); thresholds below 0.05 / at 0 use the dense path from the start.  Degenerate boxes included."""
    g = np.random.default_rng(n)
    c = g.uniform(100, 140, size=(n, 2)).astype(np.float32)           # all centres inside a 40 px window
    wh = g.uniform(30, 60, size=(n, 2)).astype(np.float32)
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    b[::97, 2] = b[::97, 0]                                            # zero width
    b[5::211, 3] = b[5::211, 1] - 1.0                                  # negative height
    s = g.uniform(0, 1, size=n).astype(np.float32)
    i = g.integers(0, groups, size=n).astype(np.int64)
    ref = O.batched_nms(b, s, i, thr)
    keep = ops.batched_nms(dev(b), dev(s), dev(i), thr).cpu().numpy()
    assert np.array_equal(keep, ref)


@pytest.mark.parametrize("kind,n,groups,thr", [("mixed_sizes", 20000, 3, 0.5), ("far_coords", 12000, 2, 0.7),
                                               ("line", 15000, 1, 0.3), ("one_big_segment", 40000, 1, 0.5),
                                               ("nonfinite", 6000, 2, 0.45), ("thr_high", 10000, 4, 0.95),
                                               ("thr_edge", 10000, 4, 0.05)])
def test_nms_adversarial_geometry(ops, kind, n, groups, thr):
    """Shapes that stress the grid pruning of the general engine: extents over four orders of magnitude, large
    coordinates (grid padding / fp32 rounding of the centres), all centres on one line (degenerate grid axis), one
    segment too large for the shared-memory resolve state, inf / NaN boxes, thresholds at both ends of the range."""
    g = np.random.default_rng(len(kind) * 1000 + n)
    c = g.uniform(0, 640, size=(n, 2))
    wh = g.uniform(8, 160, size=(n, 2))
    if kind == "mixed_sizes":
        wh = 10.0 ** g.uniform(-1, 3, size=(n, 2))
        c[: n // 4] = c[n // 4: n // 2]                     # coincident centres with unrelated extents
    elif kind == "far_coords":
        c = c + np.array([3.0e6, -7.0e5])
    elif kind == "line":
        c[:, 1] = 123.0
        wh[:, 1] = 40.0
    elif kind == "thr_high":
        c = g.uniform(0, 64, size=(n, 2))
        wh = g.uniform(30, 34, size=(n, 2))                 # near-identical boxes: IoU above 0.95 does occur
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    if kind == "nonfinite":
        b[::50, 2] = np.inf
        b[7::90, 0] = -np.inf
        b[11::130, 1] = np.nan
        b[13::170] = np.nan
    s = g.uniform(0, 1, size=n).astype(np.float32)
    m = n // 3
    s[0:3 * m:3] = s[1:3 * m:3]                             # score ties
    i = g.integers(0, groups, size=n).astype(np.int64)
    ref = O.batched_nms(b, s, i, thr)
    keep = ops.batched_nms(dev(b), dev(s), dev(i), thr).cpu().numpy()
    assert np.array_equal(keep, ref)


def test_nms_empty_and_props(ops):
    e = ops.batched_nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), 0.5)
    assert e.numel() == 0 and e.dtype == torch.int64
    b, s, i = synth.nms_boxes(8000, 4, seed=77)
    keep = ops.batched_nms(dev(b), dev(s), dev(i), 0.5)
    # idempotence: NMS of the survivors keeps every survivor, in the same order
    again = ops.batched_nms(dev(b)[keep], dev(s)[keep], dev(i)[keep], 0.5)
    assert torch.equal(again, torch.arange(keep.numel(), device="cuda"))
    assert bool((s.cuda()[keep][1:] <= s.cuda()[keep][:-1]).all())
    with pytest.raises(RuntimeError):
        ops.batched_nms(b, s, i, 0.5)  # CPU tensors are refused: no fallback


def test_nms_live_torchvision_cpu(ops):
    """torchvision is installed on the box: cross-check against its CPU kernel directly."""
    import torchvision
    b, s, i = synth.nms_boxes(12000, 5, seed=123)
    ref = torchvision.ops.boxes._batched_nms_vanilla(b, s, i, 0.55).numpy()
    keep = ops.batched_nms(dev(b), dev(s), dev(i), 0.55).cpu().numpy()
    assert np.array_equal(keep, canon(ref, s.numpy()))


# ------------------------------------------------------------------------- decode + NMS (B5, a1-a8)
def _detect_vs_oracle(ops, raws, H, W, C, og, iou, thr, allow, tracked, variant, order="image", nms_path="auto"):
    """Keep-lists must be identical to the oracle's, except for candidates that are PROVEN marginal (a score or a
    deciding IoU on the threshold to within the last-bit differences of expf; oracle/parity.py): the measured
    counts are printed (pytest -s / the captured log) and anything unexplained fails."""
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = O.decode_inference(raws, anc, H, W, og)
    ref = O.post_process(preds, iou, thr, allow, tracked)
    det = ops.detect([dev(r) for r in raws], anc, (H, W), C, og_size=og, iou_threshold=iou, score_threshold=thr,
                     box_allowance=allow, tracked_classes=tracked, order=order, variant=variant, nms_path=nms_path)
    keep = det.keep_idxs.cpu().numpy()
    rows = det.pred_boxes.cpu().numpy()
    img = det.sample_idxs.cpu().numpy()
    N = preds.shape[1]
    assert np.array_equal(img, keep // N)
    assert int(det.counts.sum()) == keep.shape[0]
    miss = np.setxor1d(keep, ref["keep"])
    par = explain_keep_mismatches(ref["score"], ref["xyxy"], N, ref["keep"], keep, iou, thr)
    print("keep-list parity B=%d %dx%d %s: %s" % (raws[0].shape[0], H, W, nms_path, summarize(par)))
    assert not par["unexplained"], summarize(par)
    if order == "global":
        assert np.all(np.diff(rows[:, 0]) <= 0)
    else:
        assert np.all(np.diff(img) >= 0)
        for b in np.unique(img):
            assert np.all(np.diff(rows[img == b, 0]) <= 0)
    if miss.size == 0:  # align rows by candidate index (scores may differ in the last ulp, so not by score order)
        a = rows[np.argsort(keep, kind="stable")]
        r = ref["pred_boxes"][np.argsort(ref["keep"], kind="stable")]
        cols = [0, 2, 3, 4, 5]
        assert_close(a[:, cols], r[:, cols], rtol=1e-5, atol=2e-5 * max(H, W), what="pred_boxes")
        # class = argmax over fp32 sigmoids, first index on ties: two logits a few 1e-7 apart can round to the same
        # sigmoid under one expf and to different ones under another (CUDA vs glibc vs ATen), so a handful of
        # near-tie rows per 100k may legitimately pick the other class; everything else must match exactly
        assert int((a[:, 1] != r[:, 1]).sum()) <= max(1, a.shape[0] // 20000), "class ids"
    return det, ref


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("name", ["post_sq64", "post_sq64_lowthr", "post_T128_tracked", "post_T128"])
def test_detect_golden(ops, name, variant):
    """Fused decode+NMS against what the unmodified reference handed to its drawing code."""
    g = golden(name)
    gd = golden(str(g["decode_case"]))
    B, H, W, C, seed, og0, og1 = (int(v) for v in gd["params"])
    raws = synth.raw_head_outputs(B, H, W, C, str(gd["dist"]), seed)
    og = None if og0 < 0 else (og0, og1)
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    det, _ = _detect_vs_oracle(ops, raws, H, W, C, og, float(g["iou"]), float(g["thr"]), allow, tracked, variant)
    counts = g["per_image_counts"]
    ref_rows = rows_canon(g["per_image"], np.repeat(np.arange(len(counts)), counts))
    got_img = np.unique(det.sample_idxs.cpu().numpy(), return_inverse=True)[1]
    got = rows_canon(det.pred_boxes.cpu().numpy(), got_img)
    assert_close(got, ref_rows, rtol=1e-5, atol=2e-5 * max(H, W), what="rows vs reference")
    assert np.array_equal(got[:, 1], ref_rows[:, 1])


@pytest.mark.parametrize("name", ["post_sq64", "post_sq64_lowthr", "post_T128_tracked", "post_T128"])
def test_post_process_golden(ops, name):
    """ops.post_process on the decoded [B, N, 85] tensor (what inference_det.post_process_preds receives) against
    the rows the unmodified reference produced, and bitwise against the fused path from the raw head outputs."""
    g = golden(name)
    gd = golden(str(g["decode_case"]))
    B, H, W, C, seed, og0, og1 = (int(v) for v in gd["params"])
    raws = synth.raw_head_outputs(B, H, W, C, str(gd["dist"]), seed)
    og = None if og0 < 0 else (og0, og1)
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = torch.cat([ops.decode_scale(dev(r), a, (H, W), True, og).reshape(B, -1, C + 5) for r, a in zip(raws, anc)], 1).contiguous()
    det = ops.post_process(preds, (H, W), C, float(g["iou"]), float(g["thr"]), allow, tracked, order="global")
    counts = g["per_image_counts"]
    ref_rows = rows_canon(g["per_image"], np.repeat(np.arange(len(counts)), counts))
    got_img = np.unique(det.sample_idxs.cpu().numpy(), return_inverse=True)[1]
    got = rows_canon(det.pred_boxes.cpu().numpy(), got_img)
    assert_close(got, ref_rows, rtol=1e-5, atol=2e-5 * max(H, W), what="rows vs reference")
    assert np.array_equal(got[:, 1], ref_rows[:, 1])
    assert bool((det.pred_boxes[1:, 0] <= det.pred_boxes[:-1, 0]).all())      # the reference's global score order
    fused = ops.detect([dev(r) for r in raws], anc, (H, W), C, og_size=og, iou_threshold=float(g["iou"]),
                       score_threshold=float(g["thr"]), box_allowance=allow, tracked_classes=tracked, order="global")
    assert torch.equal(fused.pred_boxes, det.pred_boxes) and torch.equal(fused.keep_idxs, det.keep_idxs)
    with pytest.raises(RuntimeError):
        ops.post_process(preds[:, :-1].contiguous(), (H, W), C)             # N does not match the input shape


@pytest.mark.parametrize("name", ["segpost_T128", "segpost_T128_tracked"])
@pytest.mark.parametrize("nms_path", ["auto", "general"])
def test_seg_post_process_golden(ops, name, nms_path):
    """SURVEY 8 f2, inference side: ops.post_process on rows [obj, cls*C, x,y,w,h, 4 mask coefficients] (what
    inference_seg.post_process_preds receives) against the rows and the drawn masks of the unmodified reference
    (unit protos: mask = coef > 0, which pins the gather of the extra columns), and bitwise against the 85-column call."""
    g = golden(name)
    gd = golden(str(g["decode_case"]))
    B, H, W, C, seed, og0, og1 = (int(v) for v in gd["params"])
    raws = synth.raw_head_outputs(B, H, W, C, str(gd["dist"]), seed)
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    base = torch.cat([ops.decode_scale(dev(r), a, (H, W), True, None).reshape(B, -1, C + 5) for r, a in zip(raws, anc)], 1)
    extra = seg_extra_columns(B, base.shape[1], 4, int(g["extra_seed"]))
    preds = torch.cat([base, dev(extra)], dim=-1).contiguous()
    det = ops.post_process(preds, (H, W), C, float(g["iou"]), float(g["thr"]), allow, tracked, order="global", nms_path=nms_path)
    rows, keep = det.pred_boxes.clone(), det.keep_idxs.clone()
    coefs = ops.extra_columns(preds, det, C)
    assert coefs.shape == (rows.shape[0], 4)
    counts = g["per_image_counts"]
    ref_img = np.repeat(np.arange(len(counts)), counts)
    got_img = np.unique(det.sample_idxs.cpu().numpy(), return_inverse=True)[1]
    po, pr = rows_order(rows.cpu().numpy(), got_img), rows_order(g["per_image"], ref_img)
    assert_close(rows.cpu().numpy()[po], g["per_image"][pr], rtol=1e-5, atol=2e-5 * max(H, W), what="rows vs reference")
    assert np.array_equal(coefs.cpu().numpy()[po] > 0, g["masks"][pr].astype(bool))
    plain = ops.post_process(base.contiguous(), (H, W), C, float(g["iou"]), float(g["thr"]), allow, tracked, order="global")
    assert torch.equal(plain.pred_boxes, rows) and torch.equal(plain.keep_idxs, keep)


@pytest.mark.parametrize("nms_path", ["auto", "general", "per_image_single", "per_image_lean"])
@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("B,H,W,C,dist,og,iou,thr,allow,tracked,order", [
    (2, 96, 64, 3, "N", (120, 100), 0.5, 0.2, 4, None, "image"),        # non-square, rescale, even row length (D=8)
    (1, 96, 64, 3, "N", (96, 100), 0.5, 0.2, None, None, "global"),     # `and` guard: no rescale
    (4, 320, 320, 80, "T", None, 0.65, 0.001, 4, None, "image"),
    (4, 320, 320, 80, "T", None, 0.65, 0.001, 4, None, "global"),
    (3, 320, 320, 80, "R", None, 0.65, 0.001, 4, None, "global"),       # every candidate survives (K = 6300 > per-image cap)
    (2, 160, 160, 80, "R", None, 0.65, 0.001, 4, None, "image"),        # K = 1575, every box overlaps its neighbours
    (5, 160, 160, 7, "N", None, 0.35, 0.3, 4, (1, 4), "image"),         # D = 12, ragged tiles, class filter
    (3, 256, 256, 5, "N", None, 0.2, 0.05, 4, None, "image"),           # low IoU threshold: long suppression chains
    (2, 640, 640, 80, "TP", (720, 1280), 0.35, 0.3, 4, (1, 4, 7, 16, 17), "image"),  # config 5 shape
])
def test_detect_oracle(ops, variant, nms_path, B, H, W, C, dist, og, iou, thr, allow, tracked, order):
    raws = synth.raw_head_outputs(B, H, W, C, dist, seed=7)
    _detect_vs_oracle(ops, raws, H, W, C, og, iou, thr, allow, tracked, variant, order, nms_path=nms_path)


@pytest.mark.parametrize("B,dist,iou", [(1, "T", 0.65), (100, "T", 0.65), (90, "N", 0.4)])
def test_detect_batch_extremes(ops, B, dist, iou):
    """B = 1 (video frames) and B > #SM/2 (one CTA per image, several look-back windows over the images)."""
    H = W = 128
    raws = synth.raw_head_outputs(B, H, W, 80, dist, seed=3)
    _detect_vs_oracle(ops, raws, H, W, 80, None, iou, 0.001 if dist == "T" else 0.2, 4, None, 0, "image")


def test_detect_dense_overlaps_spill(ops):
    """Every box overlaps dozens of neighbours: the overlap edges outgrow the shared-memory list and are replayed
    from the spill list; the per-image path must still agree bitwise with the general engine."""
    B, H, W, C = 3, 224, 224, 80
    raws = [dev(r) for r in synth.raw_head_outputs(B, H, W, C, "R", seed=5)]   # K = 3087 per image, all overlapping
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    a = ops.detect(raws, anc, (H, W), C, iou_threshold=0.5, score_threshold=0.001, box_allowance=4, nms_path="auto")
    assert int(a.candidates.max()) <= 4096          # stayed on the per-image path
    a = [t.clone() for t in (a.pred_boxes, a.sample_idxs, a.keep_idxs, a.counts)]
    g = ops.detect(raws, anc, (H, W), C, iou_threshold=0.5, score_threshold=0.001, box_allowance=4, nms_path="general")
    for x, y in zip(a, (g.pred_boxes, g.sample_idxs, g.keep_idxs, g.counts)):
        assert torch.equal(x, y)


def test_detect_large_per_image_variant(ops):
    """More than 4,096 survivors per image: the shim steps up to the 8,192-survivor kernel (boxes in L2) and the
    rows equal the general engine's bitwise."""
    B, H, W, C = 3, 1280, 1280, 80
    raws = [dev(r) for r in synth.raw_head_outputs(B, H, W, C, "T", seed=9)]        # ~6,800 survivors per image
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    a = ops.detect(raws, anc, (H, W), C, iou_threshold=0.65, score_threshold=0.001, box_allowance=4, nms_path="auto")
    assert 4096 < int(a.candidates.max()) <= 8192
    a = [t.clone() for t in (a.pred_boxes, a.sample_idxs, a.keep_idxs, a.counts)]
    for path in ("per_image_large", "general"):
        g = ops.detect(raws, anc, (H, W), C, iou_threshold=0.65, score_threshold=0.001, box_allowance=4, nms_path=path)
        for x, y in zip(a, (g.pred_boxes, g.sample_idxs, g.keep_idxs, g.counts)):
            assert torch.equal(x, y), path
    # small inputs through the large kernel as well
    raws = [dev(r) for r in synth.raw_head_outputs(4, 320, 320, C, "T", seed=2)]
    s1 = ops.detect(raws, anc, (320, 320), C, iou_threshold=0.5, score_threshold=0.001, box_allowance=4, nms_path="per_image")
    s1 = [t.clone() for t in (s1.pred_boxes, s1.keep_idxs)]
    s2 = ops.detect(raws, anc, (320, 320), C, iou_threshold=0.5, score_threshold=0.001, box_allowance=4, nms_path="per_image_large")
    assert torch.equal(s1[0], s2.pred_boxes) and torch.equal(s1[1], s2.keep_idxs)


def test_detect_paths_agree_exactly(ops):
    """The one-CTA-per-image NMS and the general segmented engine produce identical rows (bitwise)."""
    B, H, W, C = 8, 640, 640, 80
    raws = [dev(r) for r in synth.raw_head_outputs(B, H, W, C, "T", seed=11)]
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    for iou in (0.65, 0.3, 0.1):
        a = ops.detect(raws, anc, (H, W), C, iou_threshold=iou, score_threshold=0.001, box_allowance=4, nms_path="auto")
        a = [t.clone() for t in (a.pred_boxes, a.sample_idxs, a.keep_idxs, a.counts)]
        for path in ("general", "per_image_single", "per_image_lean"):
            g = ops.detect(raws, anc, (H, W), C, iou_threshold=iou, score_threshold=0.001, box_allowance=4, nms_path=path)
            for x, y in zip(a, (g.pred_boxes, g.sample_idxs, g.keep_idxs, g.counts)):
                assert torch.equal(x, y), (iou, path)


def test_detect_lean_kernel_steps_up(ops):
    """Throughput plans start on the lean per-image kernel (512 threads, 46 KB, up to 2,048 survivors: it shares an SM
    with the decode CTAs of other batches).  Rows are bitwise those of the general engine; an image with more survivors
    moves the plan to the 1024-thread kernel (and the configuration's hint with it), heavy overlap spills the lean
    kernel's 4,096-edge shared-memory list to the global list."""
    C = 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    dv = torch.device("cuda", 0)
    for H, dist, iou, expect_path in ((640, "T", 0.65, 5), (160, "R", 0.5, 5), (224, "R", 0.5, 2)):
        B = 3
        raws = [dev(r) for r in synth.raw_head_outputs(B, H, H, C, dist, seed=5)]
        ops._nms_path_hint.clear()
        plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (H, H), C, dv, None, iou, 0.001, 4, throughput=True)
        assert plan.params.nms_path == 5
        plan.enqueue(raws)
        d = plan.result()
        assert plan.params.nms_path == expect_path, (H, dist, int(d.candidates.max()))
        a = [t.clone() for t in (d.pred_boxes, d.sample_idxs, d.keep_idxs, d.counts)]
        again = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (H, H), C, dv, None, iou, 0.001, 4, throughput=True)
        assert again.params.nms_path == expect_path          # the hint is remembered per configuration
        g = ops.detect(raws, anc, (H, H), C, iou_threshold=iou, score_threshold=0.001, box_allowance=4, nms_path="general")
        for x, y in zip(a, (g.pred_boxes, g.sample_idxs, g.keep_idxs, g.counts)):
            assert torch.equal(x, y), (H, dist)
    ops._nms_path_hint.clear()


@pytest.mark.parametrize("depth,prio", [(1, True), (3, True), (3, False)])
def test_detect_pipeline_matches_single_stream(ops, depth, prio):
    """ops.DetectPipeline (batches in flight on several streams, own scratch each; the lean NMS kernel, on a second,
    higher-priority stream per slot when `prio`) returns, batch by batch, exactly the rows of a single-stream plan --
    also when a slot is reused and when inputs differ per batch."""
    B, H, W, C = 8, 640, 640, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    batches = [[dev(r) for r in synth.raw_head_outputs(B, H, W, C, "T", seed=20 + i)] for i in range(5)]
    shapes = [tuple(r.shape) for r in batches[0]]
    ref = []
    for raws in batches:
        d = ops.detect(raws, anc, (H, W), C, iou_threshold=0.65, score_threshold=0.001, box_allowance=4, tracked_classes=[0, 3, 17])
        ref.append([t.clone() for t in (d.pred_boxes, d.sample_idxs, d.keep_idxs, d.counts)])
    pipe = ops.DetectPipeline(shapes, anc, (H, W), C, torch.device("cuda", 0), None, 0.65, 0.001, 4, [0, 3, 17], depth=depth,
                              nms_priority=prio)
    assert (pipe.plans[0].params.nms_stream is not None) == (prio and depth > 1)
    assert pipe.plans[0].params.nms_path == (5 if depth > 1 else 0)
    slots = {}
    for i, raws in enumerate(batches):
        slot = pipe.submitted % pipe.depth
        if slot in slots:                       # take the earlier batch's rows before its slot is reused
            j = slots.pop(slot)
            got = pipe.result(slot)
            for x, y in zip(ref[j], (got.pred_boxes, got.sample_idxs, got.keep_idxs, got.counts)):
                assert torch.equal(x, y), (depth, j)
        assert pipe.submit(raws) == slot
        slots[slot] = i
    for slot, j in slots.items():
        got = pipe.result(slot)
        for x, y in zip(ref[j], (got.pred_boxes, got.sample_idxs, got.keep_idxs, got.counts)):
            assert torch.equal(x, y), (depth, j)
    pipe.join()


def test_detect_rows_on_the_host(ops):
    """DetectPlan.result_host() -- counts + rows in one pinned buffer by one copy, per-image arrays by CSR
    offsets -- gives exactly the per-image box arrays of the reference's host loop (inference_det.py:100-129), here
    taken from plan.result(); also when the optimistic copy was too small and when nothing survives."""
    B, H, W, C = 6, 320, 320, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    for dist, thr, tracked in (("TP", 0.3, [1, 4, 7, 16, 17]), ("T", 0.001, None), ("T", 0.99, None)):
        raws = [dev(r) for r in synth.raw_head_outputs(B, H, W, C, dist, seed=11)]
        plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (H, W), C, torch.device("cuda", 0), None, 0.5, thr, 4, tracked)
        for rep in range(2):          # second round: the copy size has adapted to the row count
            plan.enqueue(raws)
            d = plan.result()
            rows, img = d.pred_boxes.cpu().numpy().copy(), d.sample_idxs.cpu().numpy().copy()
            plan.enqueue(raws)
            plan.enqueue_host_copy()
            h = plan.result_host()
            assert h.rows.shape == rows.shape and np.array_equal(h.rows, rows)
            assert int(h.offsets[-1]) == rows.shape[0]
            seen = 0
            for b, boxes in h.per_image():
                assert np.array_equal(boxes, rows[img == b])
                seen += boxes.shape[0]
            assert seen == rows.shape[0]


def test_detect_config2_full_size(ops):
    """BASELINE config 2 (B=64, 640^2, conf 0.001, IoU 0.65, dist T): ALL 64 images against the oracle (the
    oracle needs ~2 s per image and runs one thread per image), plus size-independent properties."""
    B, H, W, C = 64, 640, 640, 80
    raws = synth.raw_head_outputs(B, H, W, C, "T", seed=7)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    det, ref = _detect_vs_oracle(ops, raws, H, W, C, None, 0.65, 0.001, 4, None, 0)
    keep = det.keep_idxs.cpu().numpy()
    graws = [dev(r) for r in raws]
    # both decode variants and both row orders agree exactly
    det1 = ops.detect(graws, anc, (H, W), C, iou_threshold=0.65, score_threshold=0.001, box_allowance=4, variant=1,
                      order="global")
    assert np.array_equal(np.sort(det1.keep_idxs.cpu().numpy()), np.sort(keep))
    assert bool((det1.pred_boxes[1:, 0] <= det1.pred_boxes[:-1, 0]).all())
    # idempotence: batched NMS over the kept boxes, grouped by image, keeps all of them
    det = ops.detect(graws, anc, (H, W), C, iou_threshold=0.65, score_threshold=0.001, box_allowance=4)
    again = ops.batched_nms(det.pred_boxes[:, 2:6].contiguous(), det.pred_boxes[:, 0].contiguous(), det.sample_idxs, 0.65)
    assert again.numel() == keep.shape[0]
    assert int(det.counts.sum()) == keep.shape[0] and int(det.candidates.min()) > 0


def test_detect_config4_vs_oracle(ops):
    """BASELINE config 4 inference side (1280^2, 100,800 candidates per image, dist T: ~6,800 survivors per image,
    i.e. the 8,192-survivor per-image kernel with the boxes in L2) against the oracle, which runs the reference's
    NMS over all 100,800 candidates of an image (~30 s per image, one thread per image)."""
    B, H, W, C = 2, 1280, 1280, 80
    raws = synth.raw_head_outputs(B, H, W, C, "T", seed=9)
    det, _ = _detect_vs_oracle(ops, raws, H, W, C, None, 0.65, 0.001, 4, None, 0)
    assert 4096 < int(det.candidates.max()) <= 8192


def test_detect_config5_video_frames(ops):
    """BASELINE config 5: batch-1 frames at 640^2, dist T+P re-seeded per frame, og_size (720, 1280), IoU 0.35,
    score 0.3, tracked classes: 50 frames through the B=1 plan, each against the oracle (the oracle processes the
    50 frames as one batch, one thread per frame)."""
    F, H, W, C = 50, 640, 640, 80
    tracked = list(synth.tracked_classes_default())
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    frames = [synth.raw_head_outputs(1, H, W, C, "TP", seed=7 + f) for f in range(F)]
    stacked = [torch.cat([fr[s] for fr in frames], 0) for s in range(3)]
    ref = O.post_process(O.decode_inference(stacked, anc, H, W, (720, 1280)), 0.35, 0.3, 4, tracked)
    N = synth.candidates_per_image(H, W)
    plan = ops.DetectPlan([tuple(r.shape) for r in frames[0]], anc, (H, W), C, torch.device("cuda", 0), (720, 1280), 0.35, 0.3,
                          4, tracked)
    got = []
    for f, fr in enumerate(frames):
        plan.enqueue([dev(r) for r in fr])
        d = plan.result()
        assert int(d.counts[0]) == d.keep_idxs.numel()
        got.append(d.keep_idxs.cpu().numpy() + f * N)
    got = np.concatenate(got)
    par = explain_keep_mismatches(ref["score"], ref["xyxy"], N, ref["keep"], got, 0.35, 0.3)
    print("keep-list parity config 5 (50 frames, B=1): " + summarize(par))
    assert not par["unexplained"], summarize(par)
    assert got.size > 0


def test_detect_host_result_plan(ops):
    """``DetectPlan(host_result=True)`` (bg_detect_params.host_flag): the kernels write rows and counts into page-locked host
    memory and store a sequence number last; the host polls it instead of copying and synchronising.  Rows, image
    offsets and candidate indices must be bitwise those of the ordinary plan: batch-1 frames of config 5 (the per-image
    kernel stores the flag itself), a small batch in the reference's global row order and on the general engine (the
    flag is stored by a one-thread kernel behind them)."""
    H, W, C = 640, 640, 80
    tracked = list(synth.tracked_classes_default())
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    d0 = torch.device("cuda", 0)
    frames = [[dev(r) for r in synth.raw_head_outputs(1, H, W, C, "TP", seed=7 + f)] for f in range(12)]
    shapes = [tuple(r.shape) for r in frames[0]]
    a = ops.DetectPlan(shapes, anc, (H, W), C, d0, (720, 1280), 0.35, 0.3, 4, tracked)
    b = ops.DetectPlan(shapes, anc, (H, W), C, d0, (720, 1280), 0.35, 0.3, 4, tracked, host_result=True)
    rows = 0
    for fr in frames * 3:
        a.enqueue(fr)
        da = a.result()
        b.enqueue(fr)
        hb = b.result_host()
        assert np.array_equal(hb.rows.view(np.uint32), da.pred_boxes.cpu().numpy().view(np.uint32))
        assert list(hb.offsets) == [0, int(da.counts[0])]
        b.enqueue(fr)
        db = b.result()                                     # the Detections form of the same buffers (host tensors)
        assert not db.pred_boxes.is_cuda and torch.equal(db.keep_idxs, da.keep_idxs.cpu()) and torch.equal(db.counts, da.counts)
        rows += hb.rows.shape[0]
    assert rows > 0
    B = 4
    raws = [dev(r) for r in synth.raw_head_outputs(B, 256, 256, C, "T", seed=3)]
    shapes = [tuple(r.shape) for r in raws]
    for order, path in (("image", "auto"), ("global", "auto"), ("image", "general"), ("global", "general")):
        a = ops.DetectPlan(shapes, anc, (256, 256), C, d0, None, 0.65, 0.001, 4, None, order, 0, path)
        b = ops.DetectPlan(shapes, anc, (256, 256), C, d0, None, 0.65, 0.001, 4, None, order, 0, path, host_result=True)
        for _ in range(3):
            a.enqueue(raws)
            da = a.result()
            b.enqueue(raws)
            db = b.result()
            assert da.pred_boxes.shape[0] > 0
            assert torch.equal(db.pred_boxes.view(torch.int32), da.pred_boxes.cpu().view(torch.int32)), (order, path)
            assert torch.equal(db.keep_idxs, da.keep_idxs.cpu()) and torch.equal(db.sample_idxs, da.sample_idxs.cpu())
            assert torch.equal(db.counts, da.counts) and torch.equal(db.candidates, da.candidates)
    print("host-result plan: %d rows over 36 batch-1 frames and 12 small batches, bitwise those of the ordinary plan" % rows)


@pytest.mark.parametrize("name", ["dec_sq64", "dec_rect_rescale", "dec_rect_norescale", "dec_T128"])
def test_decode_scale_golden(ops, name):
    g = golden(name)
    B, H, W, C, seed, og0, og1 = (int(v) for v in g["params"])
    raws = synth.raw_head_outputs(B, H, W, C, str(g["dist"]), seed)
    og = None if og0 < 0 else (og0, og1)
    outs = [ops.decode_scale(dev(r), synth.anchors_tensor(s), (H, W), True, og).cpu().reshape(B, -1, C + 5)
            for r, s in zip(raws, synth.SCALES)]
    preds = torch.cat(outs, 1).numpy()
    assert_close(preds[..., C + 1:], g["boxes"], rtol=1e-5, atol=2e-5 * max(H, W), what="decoded boxes")
    assert digest(np.ascontiguousarray(preds[..., :C + 1])) == str(g["logits_digest"])
    tr = ops.decode_scale(dev(raws[0]), synth.anchors_tensor("sm"), (H, W), False).cpu().numpy()
    assert_close(tr[..., C + 1:], g["train_sm_boxes"], rtol=1e-5, atol=1e-6, what="training decode")


def test_decode_scale_segmentation_rows(ops):
    """_get_scale_pred of the segmentation head (modules/detection.py:126-134,164-167): rows [obj, cls, box, K mask
    coefficients, further columns] -- the box decodes as in the detection rows (bitwise the same kernel arithmetic), the
    coefficients go through tanh (fp32, rtol 1e-5 against numpy), anything behind them is copied."""
    B, H, W, C, K, X = 2, 128, 160, 7, 8, 3
    g = torch.Generator().manual_seed(11)
    a = synth.anchors_tensor("md")
    raw = torch.randn(B, H // 16, W // 16, 3, 5 + C + K + X, generator=g) * 2.0
    out = ops.decode_scale(dev(raw), a, (H, W), True, (200, 300), num_classes=C, tanh_cols=K).cpu()
    base = ops.decode_scale(dev(raw[..., : 5 + C].contiguous()), a, (H, W), True, (200, 300)).cpu()
    assert torch.equal(out[..., : 5 + C], base)
    assert_close(out[..., 5 + C: 5 + C + K].numpy(), np.tanh(raw[..., 5 + C: 5 + C + K].numpy().astype(np.float64)), rtol=1e-5, atol=1e-7,
                 what="tanh of the mask coefficients")
    assert torch.equal(out[..., 5 + C + K:], raw[..., 5 + C + K:])
    with pytest.raises(RuntimeError):
        ops.decode_scale(dev(raw), a, (H, W), True, None, num_classes=C, tanh_cols=K + X + 1)


# -------------------------------------------------------------------------- target assignment (B1)
def _assign_check(ops, t, ny, nx, sc):
    anc = synth.anchors_tensor(sc)
    ref = O.build_target_by_scale(t, (ny, nx), anc, 4.0, 0.5)
    idx, cls, a, box, m, k = ops.build_target_by_scale(dev(t), (ny, nx), anc.cuda(), 4.0, 0.5)
    assert m is None and k is None
    assert all(x.dtype == torch.int64 for x in idx) and cls.dtype == torch.int64
    assert np.array_equal(torch.stack(idx, 0).cpu().numpy(), np.stack(ref[0], 0))
    assert np.array_equal(cls.cpu().numpy(), ref[1])
    assert np.array_equal(a.cpu().numpy(), ref[2])
    assert np.array_equal(box.cpu().numpy(), ref[3])
    return cls.numel()


@pytest.mark.parametrize("name", ["c1", "b8g100", "adv", "empty"])
def test_assign_golden(ops, name):
    g = golden("assign")
    t = {"c1": lambda: synth.targets(2, 20, 80, 0, fixed=False), "b8g100": lambda: synth.targets(8, 100, 80, 0),
         "adv": lambda: synth.adversarial_targets(2, 80), "empty": lambda: torch.zeros(0, 6)}[name]()
    for key in sorted(k[:-4] for k in g.files if k.endswith("_idx") and k.startswith(name + "_")):
        _, fm, sc = key.rsplit("_", 2)
        ny, nx = (int(v) for v in fm.split("x"))
        idx, cls, a, box, _, _ = ops.build_target_by_scale(dev(t), (ny, nx), synth.anchors_tensor(sc), 4.0, 0.5)
        got = torch.stack(idx, 0).cpu().numpy() if cls.numel() else np.zeros((4, 0), np.int64)
        assert np.array_equal(got, g[key + "_idx"]), key
        assert np.array_equal(cls.cpu().numpy(), g[key + "_cls"])
        assert np.array_equal(a.cpu().numpy().reshape(-1, 2), g[key + "_anc"])
        assert np.array_equal(box.cpu().numpy().reshape(-1, 4), g[key + "_box"])


@pytest.mark.parametrize("name", ASSIGN_VARIANTS)
def test_assign_variants_golden(ops, name):
    """Segmentation (overlap_masks) and keypoint-column variants: bit-exact with the reference's outputs."""
    g = golden("assign_variants")
    t, overlap, bs = assign_variant_case(name)
    for (ny, nx), sc in zip(((16, 16), (8, 8), (4, 4)), synth.SCALES):
        idx, cls, a, box, tm, kp = ops.build_target_by_scale(dev(t), (ny, nx), synth.anchors_tensor(sc), 4.0, 0.5, overlap, bs)
        k = f"{name}_{sc}"
        assert np.array_equal(torch.stack(idx, 0).cpu().numpy(), g[k + "_idx"]) and np.array_equal(cls.cpu().numpy(), g[k + "_cls"])
        assert np.array_equal(a.cpu().numpy(), g[k + "_anc"]) and np.array_equal(box.cpu().numpy(), g[k + "_box"])
        assert (tm is None) == (k + "_tmask" not in g.files) and (kp is None) == (k + "_kpts" not in g.files)
        if tm is not None:
            assert tm.dtype == torch.int64 and np.array_equal(tm.cpu().numpy(), g[k + "_tmask"])
        if kp is not None:
            assert np.array_equal(kp.cpu().numpy(), g[k + "_kpts"])


def test_assign_variants_errors(ops):
    t = synth.targets(2, 5, 80, 0)
    with pytest.raises(ValueError):   # the reference's own error: overlap_masks=True needs batch_size
        ops.build_target_by_scale(dev(t), (8, 8), synth.anchors_tensor("md"), 4.0, 0.5, True, None)
    with pytest.raises(RuntimeError):  # image ids outside 0..batch_size-1: the reference's torch.cat raises
        ops.build_target_by_scale(dev(t), (8, 8), synth.anchors_tensor("md"), 4.0, 0.5, True, 1)


def test_assign_config3_and_4(ops):
    t3 = synth.targets(256, 100, 80, 0)          # config 3: nt = 25 600
    total = sum(_assign_check(ops, t3, s, s, sc) for s, sc in zip((80, 40, 20), synth.SCALES))
    assert total > 400000
    t4 = synth.targets(32, 300, 80, 0)           # config 4: 1280^2, 300 gt / image
    for s, sc in zip((160, 80, 40), synth.SCALES):
        _assign_check(ops, t4, s, s, sc)


# -------------------------------------------------------------------------------------- CIoU (B2)
def test_ciou(ops):
    g = golden("ciou")
    p = dev(g["p"]).requires_grad_(True)
    c = ops.compute_ciou(p, dev(g["t"]))
    (c * dev(g["w"])).sum().backward()
    assert_close(c.detach().cpu().numpy(), g["ciou"], rtol=1e-5, atol=1e-6, what="ciou vs reference")
    assert_close(p.grad.cpu().numpy(), g["grad"], rtol=1e-4, atol=1e-5, what="ciou grad vs reference autograd")
    oc, og = O.compute_ciou(g["p"], g["t"], with_grad=True)
    assert_close(c.detach().cpu().numpy(), oc, rtol=1e-5, atol=1e-6, what="ciou vs oracle")
    # the broadcasting form (detection_loss.py:231-234): targets [m, 4] against preds [m, A, 4]
    m, A = g["p"].shape[0] // 5, 5
    pb = dev(g["p"][: m * A]).reshape(m, A, 4).clone().requires_grad_(True)
    tb = dev(g["t"][:m])
    cb = ops.compute_ciou(pb, tb)
    assert tuple(cb.shape) == (m, A)
    cb.sum().backward()
    te = np.repeat(g["t"][:m], A, axis=0)
    oc2, og2 = O.compute_ciou(g["p"][: m * A], te, with_grad=True)
    assert_close(cb.detach().cpu().numpy().reshape(-1), oc2, rtol=1e-5, atol=1e-6, what="broadcast ciou vs oracle")
    assert_close(pb.grad.cpu().numpy().reshape(-1, 4), og2, rtol=1e-4, atol=1e-5, what="broadcast ciou grad vs oracle")


# -------------------------------------------------------------------------------------- loss (B3)
def _loss_case(ops, B, H, W, C, t, preds, rtol_grad=1e-4):
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    gp = [dev(p).requires_grad_(True) for p in preds]
    loss, metrics = ops.detection_loss(gp, dev(t) if t.numel() else torch.zeros(0, 6).cuda(), anc, synth.LOSS_CONFIG)
    loss.backward()
    return loss, metrics, [p.grad.cpu().numpy() for p in gp]


@pytest.mark.parametrize("name", ["loss_sq64", "loss_collide", "loss_c3_rect", "loss_empty", "loss_c1_640"])
def test_loss_golden(ops, name):
    g = golden(name)
    B, H, W, C, G, fixed, ts, ps = (int(v) for v in g["params"])
    t = synth.targets(B, G, C, ts, bool(fixed)) if G > 0 else torch.zeros(0, 6)
    preds = synth.train_preds(B, H, W, C, ps)
    loss, metrics, grads = _loss_case(ops, B, H, W, C, t, preds)
    assert_close(float(loss), float(g["loss"]), rtol=1e-5, atol=0, what="loss vs reference")
    ref_m = dict(zip((str(k) for k in g["metric_keys"]), g["metric_vals"]))
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for sc, gr in zip(synth.SCALES, grads):
        if "grad_" + sc in g.files:
            assert_close(gr, g["grad_" + sc], rtol=1e-4, atol=1e-7, what="grad " + sc)
        else:
            assert_close(gr[..., 0].astype(np.float64).sum(), float(g["grad_" + sc + "_obj_sum"]), rtol=1e-4, atol=1e-7)
            assert_close(np.abs(gr.astype(np.float64)).sum(), float(g["grad_" + sc + "_abs_sum"]), rtol=1e-4)
            ix = g["grad_" + sc + "_rows_idx"]
            assert_close(gr[ix[:, 0], ix[:, 1], ix[:, 2], ix[:, 3]], g["grad_" + sc + "_rows"], rtol=1e-4, atol=1e-7)


def test_loss_config3_shard(ops):
    """Config 3 per-GPU shard at P=8 (B=32, 100 gt/img) against the oracle, incl. the dense gradient."""
    B, H, W, C = 32, 640, 640, 80
    t = synth.targets(B, 100, C, 0)
    preds = synth.train_preds(B, H, W, C, 1)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ref_loss, ref_m, ref_g, Ms = O.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_grad=True)
    loss, metrics, grads = _loss_case(ops, B, H, W, C, t, preds)
    assert_close(float(loss), ref_loss, rtol=1e-5, what="loss")
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for a, b in zip(grads, ref_g):
        assert_close(a, b, rtol=1e-4, atol=1e-9, what="grad")


def _split(p, C):
    return (p[..., 0].contiguous(), p[..., 1:1 + C].contiguous(), p[..., 1 + C:].contiguous())


@pytest.mark.parametrize("name", ["lossraw_sq64", "lossraw_collide", "lossraw_empty", "lossraw_c1_640"])
def test_loss_raw_and_split_golden(ops, name):
    """the loss from the head's own logits -- interleaved rows (`raw`) and the head's three conv
    outputs (`split`) -- against the UNMODIFIED reference running _get_scale_pred(inference=False) + DetectionLoss on
    the same logits, gradients with respect to the logits (tests/golden/lossraw_*.npz)."""
    g = golden(name)
    B, H, W, C, G, fixed, ts, ps = (int(v) for v in g["params"])
    t = synth.targets(B, G, C, ts, bool(fixed)) if G > 0 else torch.zeros(0, 6)
    raws = synth.train_preds(B, H, W, C, ps)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    td = dev(t) if t.numel() else torch.zeros(0, 6).cuda()
    gp = [dev(p).requires_grad_(True) for p in raws]
    loss, metrics = ops.detection_loss(gp, td, anc, synth.LOSS_CONFIG, input_form="raw")
    loss.backward()
    assert_close(float(loss), float(g["loss"]), rtol=1e-5, atol=0, what="loss vs reference")
    for k, v in dict(zip((str(k) for k in g["metric_keys"]), g["metric_vals"])).items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    grads = [p.grad.cpu().numpy() for p in gp]
    for sc, gr in zip(synth.SCALES, grads):
        if "grad_" + sc in g.files:
            assert_close(gr, g["grad_" + sc], rtol=1e-4, atol=1e-7, what="grad " + sc)
        else:
            assert_close(gr[..., 0].astype(np.float64).sum(), float(g["grad_" + sc + "_obj_sum"]), rtol=1e-4, atol=1e-7)
            assert_close(np.abs(gr.astype(np.float64)).sum(), float(g["grad_" + sc + "_abs_sum"]), rtol=1e-4)
            ix = g["grad_" + sc + "_rows_idx"]
            assert_close(gr[ix[:, 0], ix[:, 1], ix[:, 2], ix[:, 3]], g["grad_" + sc + "_rows"], rtol=1e-4, atol=1e-7)
    # the split form: same arithmetic on the three column groups -> identical loss and gradients, piece by piece
    tri = [tuple(x.requires_grad_(True) for x in _split(dev(p), C)) for p in raws]
    loss_s, metrics_s = ops.detection_loss(tri, td, anc, synth.LOSS_CONFIG, input_form="split")
    loss_s.backward()
    assert_close(float(loss_s), float(loss), rtol=1e-6, atol=0, what="split form vs raw form")   # (the dense pass sums in another order)
    for k in metrics:
        assert_close(metrics_s[k], metrics[k], rtol=1e-6, atol=1e-9, what=k)
    for gr, (c, k, b) in zip(grads, tri):
        assert np.array_equal(gr[..., 0], c.grad.cpu().numpy())
        assert np.array_equal(gr[..., 1:1 + C], k.grad.cpu().numpy())
        assert np.array_equal(gr[..., 1 + C:], b.grad.cpu().numpy())
    # and the decoded form fed with the reference's own decode of the logits gives the same loss
    dec = [ops.decode_scale(dev(p), a, (H, W), False) for p, a in zip(raws, anc)]
    loss_d, _ = ops.detection_loss(dec, td, anc, synth.LOSS_CONFIG, with_metrics=False)
    assert_close(float(loss_d), float(loss), rtol=1e-6, atol=0, what="decoded form vs raw form")


def test_loss_forms_vs_oracle_config3_shard(ops):
    """The raw and split forms at the per-GPU shard of config 3 (B=32, 100 gt/img) against the oracle
    (decode + loss + chain rule), incl. the dense gradient."""
    B, H, W, C = 32, 640, 640, 80
    t = synth.targets(B, 100, C, 0)
    raws = synth.train_preds(B, H, W, C, 1)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ref_loss, ref_m, ref_g, _ = O.detection_loss(raws, t, anc, synth.LOSS_CONFIG, with_grad=True, input_form="raw")
    gp = [dev(p).requires_grad_(True) for p in raws]
    loss, metrics = ops.detection_loss(gp, dev(t), anc, synth.LOSS_CONFIG, input_form="raw")
    loss.backward()
    assert_close(float(loss), ref_loss, rtol=1e-5, what="loss (raw)")
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for a, b in zip(gp, ref_g):
        assert_close(a.grad.cpu().numpy(), b, rtol=1e-4, atol=1e-9, what="grad (raw)")
    tri = [tuple(x.requires_grad_(True) for x in _split(dev(p), C)) for p in raws]
    loss_s, _ = ops.detection_loss(tri, dev(t), anc, synth.LOSS_CONFIG, input_form="split", with_metrics=False)
    loss_s.backward()
    assert_close(float(loss_s), float(loss), rtol=1e-6, atol=0, what="split form vs raw form")
    for a, (c, k, b) in zip(gp, tri):
        assert torch.equal(a.grad[..., 0], c.grad) and torch.equal(a.grad[..., 1:1 + C], k.grad) and torch.equal(a.grad[..., 1 + C:], b.grad)


def test_loss_config4_vs_oracle(ops):
    """BASELINE config 4 training side at full size: B=32 at 1280^2 (100,800 cells/img), 300 gt/img, against the
    oracle incl. the dense gradient (1.1 GB per tensor set)."""
    B, H, W, C = 32, 1280, 1280, 80
    t = synth.targets(B, 300, C, 0)
    preds = synth.train_preds(B, H, W, C, 1)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ref_loss, ref_m, ref_g, Ms = O.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_grad=True)
    loss, metrics, grads = _loss_case(ops, B, H, W, C, t, preds)
    assert_close(float(loss), ref_loss, rtol=1e-5, what="loss")
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for a, b in zip(grads, ref_g):
        assert_close(a, b, rtol=1e-4, atol=1e-9, what="grad")
    assert sum(Ms) > 150000


def test_loss_forwards_may_precede_their_backwards(ops):
    """Two forwards, then their two backwards (gradient accumulation, several loss modules, a validation loss in
    between): every forward owns the state its backward reads.  Each gradient must equal the one obtained alone."""
    B, H, W, C = 4, 128, 128, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ta, tb = dev(synth.targets(B, 12, C, 0)), dev(synth.targets(B, 7, C, 5, fixed=False))
    pa = [dev(p).requires_grad_(True) for p in synth.train_preds(B, H, W, C, 1)]
    pb = [dev(p).requires_grad_(True) for p in synth.train_preds(B, H, W, C, 2)]

    def alone(p, t):
        q = [x.detach().clone().requires_grad_(True) for x in p]
        l, _ = ops.detection_loss(q, t, anc, synth.LOSS_CONFIG, with_metrics=False)
        l.backward()
        return float(l), [x.grad.clone() for x in q]

    la_ref, ga_ref = alone(pa, ta)
    lb_ref, gb_ref = alone(pb, tb)
    la, _ = ops.detection_loss(pa, ta, anc, synth.LOSS_CONFIG, with_metrics=False)
    lb, _ = ops.detection_loss(pb, tb, anc, synth.LOSS_CONFIG, with_metrics=False)      # second forward before the first backward
    with torch.no_grad():
        ops.detection_loss([x.detach() for x in pb], ta, anc, synth.LOSS_CONFIG)        # an evaluation loss in between
    (la + 2.0 * lb).backward()
    assert float(la) == la_ref and float(lb) == lb_ref
    for x, r in zip(pa, ga_ref):
        assert torch.equal(x.grad, r)
    for x, r in zip(pb, gb_ref):
        assert_close(x.grad.cpu().numpy(), (2.0 * r).cpu().numpy(), rtol=1e-6, atol=0, what="scaled upstream gradient")
    oracle_loss, _, og, _ = O.detection_loss([p.detach().cpu() for p in pa], ta.cpu(), anc, synth.LOSS_CONFIG, with_grad=True)
    assert_close(la_ref, oracle_loss, rtol=1e-5, what="loss vs oracle")
    for x, r in zip(pa, og):
        assert_close(x.grad.cpu().numpy(), r, rtol=1e-4, atol=1e-9, what="grad vs oracle")


def test_loss_step_graph_replays_the_eager_step(ops):
    """ops.LossStepGraph: forward + backward captured in a CUDA graph give, replay after replay and after new values
    are written into the captured tensors, exactly the eager loss and gradients."""
    B, H, W, C = 4, 128, 128, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    cfg = dict(synth.LOSS_CONFIG, num_classes=C)
    t = dev(synth.targets(B, 12, C, 0))
    stat = [dev(p).requires_grad_(True) for p in synth.train_preds(B, H, W, C, 1)]
    cells = [x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] for x in stat]
    gs = ops.LossStepGraph(stat, t, anc, cfg, input_form="raw", cells=cells)
    for seed in (1, 2, 2, 3):
        new = [dev(p) for p in synth.train_preds(B, H, W, C, seed)]
        with torch.no_grad():
            for s_, n_ in zip(stat, new):
                s_.copy_(n_)
        loss_g = float(gs.replay())
        eager = [n_.clone().requires_grad_(True) for n_ in new]
        loss_e, _ = ops.detection_loss(eager, t, anc, cfg, with_metrics=False, input_form="raw")
        loss_e.backward()
        assert loss_g == float(loss_e)
        assert_close(float(gs.combined), loss_g, rtol=1e-6, atol=0, what="combined loss, one rank")
        for s_, e_ in zip(stat, eager):
            assert torch.equal(s_.grad, e_.grad)


def test_loss_split_gradients_cleared_ahead(ops):
    """bg_loss_clear_grads + BG_LOSS_BWD_PRECLEARED (ops.PRECLEAR_SPLIT_GRADS): the class / box gradient planes cleared on
    a second stream next to the forward give the gradients of the default path bit for bit, eagerly and from a graph."""
    B, H, W, C = 4, 128, 128, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    cfg = dict(synth.LOSS_CONFIG, num_classes=C)
    t = dev(synth.targets(B, 12, C, 0))
    raws = synth.train_preds(B, H, W, C, 1)

    def run(graph):
        tri = [tuple(x.requires_grad_(True) for x in _split(dev(p), C)) for p in raws]
        if graph:
            gs = ops.LossStepGraph(tri, t, anc, cfg, input_form="split")
            loss = gs.replay()
        else:
            loss, _ = ops.detection_loss(tri, t, anc, cfg, with_metrics=False, input_form="split")
            loss.backward()
        torch.cuda.synchronize()
        return float(loss), [x.grad.clone() for p in tri for x in p]

    base = run(False)
    keep = ops.PRECLEAR_SPLIT_GRADS
    try:
        for mode in (True, "priority", False):
            ops.PRECLEAR_SPLIT_GRADS = mode
            for graph in (False, True):
                got = run(graph)
                assert got[0] == base[0], (mode, graph)
                assert all(torch.equal(a, b) for a, b in zip(got[1], base[1])), (mode, graph)
    finally:
        ops.PRECLEAR_SPLIT_GRADS = keep


def test_loss_rejects_out_of_range_ids(ops):
    """Image ids outside 0..B-1 and class ids outside 0..C-1: the reference raises IndexError (preds[batch_idx, ...],
    t_cls[range, classes]); the CUDA path drops those rows on the device -- no out-of-bounds access -- and raises
    the IndexError with the metrics read."""
    B, H, W, C = 2, 128, 128, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = [dev(p).requires_grad_(True) for p in synth.train_preds(B, H, W, C, 1)]
    good = synth.targets(B, 9, C, 0)
    for col, val in ((0, float(B)), (0, 57.0), (1, float(C)), (1, 1.0e6)):
        bad = good.clone()
        bad[3, col] = val
        with pytest.raises(IndexError):
            ops.detection_loss(preds, dev(bad), anc, synth.LOSS_CONFIG)
        loss, _ = ops.detection_loss(preds, dev(bad), anc, synth.LOSS_CONFIG, with_metrics=False)
        loss.backward()                                          # nothing is corrupted: the bad row is simply absent
        keep = torch.ones(bad.shape[0], dtype=torch.bool)
        keep[3] = False
        ref, _ = ops.detection_loss([p.detach() for p in preds], dev(good[keep]), anc, synth.LOSS_CONFIG, with_metrics=False)
        assert_close(float(loss), float(ref), rtol=1e-6, atol=0, what="loss without the bad row")
    assert all(bool(torch.isfinite(p.grad).all()) for p in preds)


def test_ratio_metrics_is_deterministic(ops):
    g = np.random.default_rng(3)
    wh = torch.from_numpy(g.uniform(0.01, 0.6, size=(300000, 2)).astype(np.float32)).cuda()
    anc = torch.tensor(sum((synth.ANCHORS[s] for s in synth.SCALES), []), dtype=torch.float32)
    first = ops.ratio_metrics_w_extras(anc, wh, 4.0)
    assert all(ops.ratio_metrics_w_extras(anc, wh, 4.0) == first for _ in range(5))
    ref = O.ratio_metrics(anc, wh.cpu(), 4.0)
    assert_close(np.array(first), np.array(ref), rtol=1e-5)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_follow_the_tensors_device(ops):
    """The reference's DDP path addresses ranks as cuda:k without calling torch.cuda.set_device: every operator must
    run on the device (and stream) of its tensors, whatever the current device is."""
    assert torch.cuda.current_device() == 0
    d1 = torch.device("cuda", 1)
    B, H, W, C = 2, 128, 128, 80
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    t = synth.targets(B, 9, C, 0)
    preds = synth.train_preds(B, H, W, C, 1)
    p0 = [p.cuda(0).requires_grad_(True) for p in preds]
    p1 = [p.to(d1).requires_grad_(True) for p in preds]
    l0, m0 = ops.detection_loss(p0, t.cuda(0), anc, synth.LOSS_CONFIG)
    l1, m1 = ops.detection_loss(p1, t.to(d1), anc, synth.LOSS_CONFIG)
    l0.backward()
    l1.backward()
    assert l1.device == d1 and float(l0) == float(l1) and m0 == m1
    for a, b in zip(p0, p1):
        assert b.grad.device == d1 and torch.equal(a.grad.cpu(), b.grad.cpu())
    raws = synth.raw_head_outputs(B, H, W, C, "TP", seed=7)
    a = ops.detect([r.cuda(0) for r in raws], anc, (H, W), C, iou_threshold=0.5, score_threshold=0.01, box_allowance=4)
    b = ops.detect([r.to(d1) for r in raws], anc, (H, W), C, iou_threshold=0.5, score_threshold=0.01, box_allowance=4)
    assert b.pred_boxes.device == d1 and torch.equal(a.pred_boxes.cpu(), b.pred_boxes.cpu()) and torch.equal(a.keep_idxs.cpu(), b.keep_idxs.cpu())
    bx, sx, ix = synth.nms_boxes(3000, 3, seed=5)
    assert torch.equal(ops.batched_nms(bx.cuda(0), sx.cuda(0), ix.cuda(0), 0.5).cpu(), ops.batched_nms(bx.to(d1), sx.to(d1), ix.to(d1), 0.5).cpu())
    idx0 = ops.build_target_by_scale(t.cuda(0), (16, 16), anc[0])
    idx1 = ops.build_target_by_scale(t.to(d1), (16, 16), anc[0])
    assert idx1[1].device == d1 and torch.equal(torch.stack(idx0[0]).cpu(), torch.stack(idx1[0]).cpu())
    # the operators added late in round 2: mask assembly, the segmentation head's decode, the host-result plan
    gq = torch.Generator().manual_seed(3)
    coefs, protos, cnt = torch.tanh(torch.randn(5, 8, generator=gq)), torch.randn(2, 8, 16, 16, generator=gq), torch.tensor([2, 3])
    m0 = ops.seg_masks(coefs.cuda(0), cnt, protos.cuda(0), (40, 48))
    m1 = ops.seg_masks(coefs.to(d1), cnt, protos.to(d1), (40, 48))
    assert m1.device == d1 and torch.equal(m0.cpu(), m1.cpu())
    rs = torch.randn(2, 8, 8, 3, 5 + 7 + 4, generator=gq)
    s0 = ops.decode_scale(rs.cuda(0), anc[1], (128, 128), True, None, num_classes=7, tanh_cols=4)
    s1 = ops.decode_scale(rs.to(d1), anc[1], (128, 128), True, None, num_classes=7, tanh_cols=4)
    assert s1.device == d1 and torch.equal(s0.cpu(), s1.cpu())
    one = [r[:1].contiguous() for r in raws]
    shapes = [tuple(r.shape) for r in one]
    h0 = ops.DetectPlan(shapes, anc, (H, W), C, torch.device("cuda", 0), None, 0.5, 0.01, 4, host_result=True)
    h1 = ops.DetectPlan(shapes, anc, (H, W), C, d1, None, 0.5, 0.01, 4, host_result=True)
    h0.enqueue([r.cuda(0) for r in one])
    h1.enqueue([r.to(d1) for r in one])
    r0, r1 = h0.result_host(), h1.result_host()
    assert r0.rows.shape[0] > 0 and np.array_equal(r0.rows.view(np.uint32), r1.rows.view(np.uint32))
    assert torch.cuda.current_device() == 0
    with pytest.raises(RuntimeError, match="one device"):
        ops.detection_loss(p0, t.to(d1), anc, synth.LOSS_CONFIG)


def test_loss_config3_full_size_properties(ops):
    """BASELINE config 3 at full size (B = 256, 100 gt/img): size-independent properties instead of the oracle.
    (1) image sharding: the 8 per-shard scalar blocks combine (shard.allreduce_loss_terms, the single all-reduce
    of the multi-GPU path) to the big-batch loss; (2) the dense gradient is finite, its objectness column is
    non-zero everywhere, the class/box columns are non-zero exactly on matched rows; (3) determinism of the
    forward; (4) assignment counts add up across shards."""
    from vision_conglomerate_b200 import shard
    B, H, W, C, G, P = 256, 640, 640, 80, 100, 8
    t = synth.targets(B, G, C, 0).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    preds = [torch.randn(B, ny, nx, 3, 5 + C, generator=g, device="cuda").requires_grad_(True) for ny, nx in synth.fmap_shapes(H, W)]
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    cfg = dict(synth.LOSS_CONFIG, num_classes=C)
    loss, _, sc_full = ops.detection_loss(preds, t, anc, cfg, with_metrics=False, return_scalars=True)
    sc_full = sc_full.clone()
    loss.backward()
    loss2, _, _ = ops.detection_loss(preds, t, anc, cfg, with_metrics=False, return_scalars=True)
    assert float(loss2) == float(loss)                                              # (3)
    # (1) + (4)
    parts, Ms = [], torch.zeros(3, dtype=torch.float64)
    cells = None
    for r in range(P):
        s, e = shard.shard_range(B, P, r)
        tl = shard.shard_targets(t, s, e)
        pl = [p.detach()[s:e].contiguous() for p in preds]
        _, _, sc = ops.detection_loss(pl, tl, anc, cfg, with_metrics=False, return_scalars=True)
        parts.append(sc.clone())
        Ms += sc[:, 6].cpu()
        cells = [p.shape[0] * p.shape[1] * p.shape[2] * p.shape[3] for p in pl]
    assert torch.equal(Ms, sc_full[:, 6].cpu())
    # emulate the all-reduce: sum of the packed per-rank terms (no process group -> allreduce_loss_terms skips it)
    M = torch.stack([p[:, 6] for p in parts]).sum(0)
    c = torch.tensor([cc * P for cc in cells], dtype=torch.float64, device="cuda")
    lbox = torch.stack([p[:, 0] * p[:, 6] for p in parts]).sum(0) / M
    lconf = torch.stack([p[:, 1] * (c / P) for p in parts]).sum(0) / c
    lcls = torch.stack([p[:, 2] * p[:, 6] for p in parts]).sum(0) / M
    sw = torch.tensor(cfg["scale_w"], dtype=torch.float64, device="cuda")
    combined = cfg["box_w"] * (sw * lbox).sum() + cfg["conf_w"] * (sw * lconf).sum() + cfg["class_w"] * (sw * lcls).sum()
    assert_close(float(combined), float(loss), rtol=1e-6, atol=0, what="sharded loss == big-batch loss")
    one = shard.allreduce_loss_terms(sc_full, [cc * P for cc in cells], cfg)        # world size 1: identity
    assert_close(float(one), float(loss), rtol=1e-6, atol=0, what="allreduce_loss_terms, single rank")
    # (2)
    for p in preds:
        gr = p.grad
        assert bool(torch.isfinite(gr).all())
        assert int((gr[..., 0] == 0).sum()) == 0
        row_nz = (gr[..., 1:] != 0).any(-1)
        assert 0 < int(row_nz.sum()) < row_nz.numel()
    matched = sum(int((p.grad[..., 1:] != 0).any(-1).sum()) for p in preds)
    assert matched <= int(sc_full[:, 6].sum())          # matched cells <= matches (duplicates share a cell)
    assert matched >= 0.8 * int(sc_full[:, 6].sum())


# ------------------------------------------------------------------------------ drop-in glue on the GPU
def test_dropin_glue_with_reference_shaped_classes(ops):
    """The reference cannot travel to the GPU box, so the drop-in layer is exercised here on stand-ins that carry
    exactly the attributes and call signatures the reference classes have (modules/detection_loss.py:40-122,
    dataset/detection_dataset.py:90-107, modules/detection.py:98-104): install() must route their call sites to
    the CUDA operators and uninstall() must restore them."""
    import torchvision
    from vision_conglomerate_b200 import dropin

    class DetectionDataset:
        @staticmethod
        def build_target_by_scale(targets, fmap_shape, anchors, anchor_threshold=4.0, edge_threshold=0.5,
                                  overlap_masks=None, batch_size=None):
            raise AssertionError("reference implementation called")

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.num_classes, self.num_keypoints = 80, None
            for k in synth.SCALES:
                setattr(self, k + "_anchors", torch.nn.Parameter(synth.anchors_tensor(k).cuda()))

    class DetectionLoss(torch.nn.Module):
        def __init__(self, model, **kw):
            super().__init__()
            self.model = model
            for k, v in kw.items():
                setattr(self, k, v)
            self.scale_w = kw.get("scale_w") or [4.0, 2.0, 1.0]

        def forward(self, preds, targets):
            raise AssertionError("reference implementation called")

        def loss_fn(self, preds, targets, anchors):
            raise AssertionError("reference implementation called")

        @staticmethod
        def compute_ciou(preds_xywh, targets_xywh, e=1e-7):
            raise AssertionError("reference implementation called")

    class DetectionNet(torch.nn.Module):
        num_keypoints = None

        def _get_scale_pred(self, scale_pred, anchors, input_shape, inference=False):
            raise AssertionError("reference implementation called")

    orig_nms = torchvision.ops.batched_nms
    dropin.install(DetectionDataset, DetectionLoss, DetectionNet)
    try:
        B, H, W, C = 2, 128, 128, 80
        t = synth.targets(B, 9, C, 0).cuda()
        preds = [p.cuda().requires_grad_(True) for p in synth.train_preds(B, H, W, C, 1)]
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        loss_mod = DetectionLoss(Model(), **synth.LOSS_CONFIG)
        loss, metrics = loss_mod(tuple(preds), t)
        loss.backward()
        ref_loss, ref_metrics = ops.detection_loss([p.detach() for p in preds], t, anc, synth.LOSS_CONFIG)
        assert float(loss) == float(ref_loss) and set(metrics) == set(ref_metrics) and len(metrics) == 10
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in preds)
        out = DetectionDataset.build_target_by_scale(t, (16, 16), anc[0].cuda())
        exp = ops.build_target_by_scale(t, (16, 16), anc[0])
        assert torch.equal(torch.stack(out[0]), torch.stack(exp[0])) and out[4] is None and out[5] is None
        p4, t4 = torch.rand(64, 4).cuda() + 0.1, torch.rand(64, 4).cuda() + 0.1
        assert torch.equal(DetectionLoss.compute_ciou(p4, t4), ops.compute_ciou(p4, t4))
        x = synth.raw_head_outputs(1, 64, 64, C, "N", 3)[0].cuda()
        dec = DetectionNet()._get_scale_pred(x, anc[0].cuda(), (64, 64), inference=True)
        assert torch.equal(dec, ops.decode_scale(x, anc[0], (64, 64), True))
        b, s, i = synth.nms_boxes(5000, 3, seed=5)
        keep = torchvision.ops.batched_nms(b.cuda(), s.cuda(), i.cuda(), 0.5)
        assert np.array_equal(keep.cpu().numpy(), canon(orig_nms(b, s, i, 0.5).numpy(), s.numpy()))
        with pytest.raises(RuntimeError, match="CUDA"):
            torchvision.ops.batched_nms(b, s, i, 0.5)
    finally:
        dropin.uninstall()
    assert torchvision.ops.batched_nms is orig_nms


# ------------------------------------------------------------------------------ anchor metrics (a13)
def test_ratio_metrics(ops):
    g = golden("ratio")
    s = ops.ratio_metrics_w_extras(g["anchors"], dev(g["wh"]), 4.0)
    assert_close(np.array(s), g["extras"], rtol=1e-5)
    assert_close(ops.ratio_metrics(g["anchors"], dev(g["wh"]), 4.0), float(g["score"]), rtol=1e-5)
    assert_close(np.array(ops.ratio_metrics_w_extras(g["anchors"], dev(g["wh"]) * 3.0, 2.0)), g["extras_x3_t2"], rtol=1e-5)


# ------------------------------------------------------------------------------------------------ f2: SegmentationLoss
def _seg_case(ops, preds, protos, t, masks, C, K, cfg):
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    gp = [dev(p).requires_grad_(True) for p in preds]
    pr = dev(protos).requires_grad_(True)
    loss, metrics = ops.segmentation_loss(gp, dev(t), pr, dev(masks), anc, cfg, C, K)
    loss.backward()
    return loss, metrics, [p.grad.cpu().numpy() for p in gp], pr.grad.cpu().numpy()


@pytest.mark.parametrize("name", ["segloss_sq64", "segloss_rect", "segloss_128", "segloss_empty_img"])
def test_segmentation_loss_golden(ops, name):
    """SegmentationLoss.forward + backward on the CUDA path (fused detection terms + the mask-term kernels)
    against the UNMODIFIED reference (modules/segmentation_loss.py:26-231, overlap_masks=True): loss rtol 1e-5, the twelve
    metrics, gradients with respect to the three prediction tensors and the protos rtol 1e-4."""
    from tests.util import seg_loss_case
    g = golden(name)
    preds, protos, t, masks, C, K = seg_loss_case(name, g)
    cfg = dict(synth.LOSS_CONFIG, seg_w=1.0)
    loss, metrics, grads, gpr = _seg_case(ops, preds, protos, t, masks, C, K, cfg)
    assert_close(float(loss), float(g["loss"]), rtol=1e-5, atol=0, what="loss vs reference")
    ref_m = dict(zip((str(k) for k in g["metric_keys"]), g["metric_vals"]))
    assert set(metrics) == set(ref_m)
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for sc, gr in zip(synth.SCALES, grads):
        assert_close(gr, g["grad_" + sc], rtol=1e-4, atol=1e-7, what="grad " + sc)
    assert_close(gpr, g["grad_protos"], rtol=1e-4, atol=1e-8, what="grad protos")


@pytest.mark.parametrize("B,S,C,K,G,mdiv,seg_w", [(4, 256, 80, 32, 10, 1, 1.0), (5, 160, 7, 16, 40, 2, 0.7), (16, 320, 80, 32, 20, 1, 1.0)])
def test_segmentation_loss_vs_oracle(ops, B, S, C, K, G, mdiv, seg_w):
    """Larger shapes against the oracle (oracle/seg_oracle.py, itself pinned to the reference by the fixtures): several
    match groups per image (more than 32 matches), pixel counts that are not multiples of the tile, 16 coefficients, a
    weight other than 1, duplicate cells."""
    from oracle import seg_oracle as SO
    preds, protos, t, masks = synth.seg_inputs(B, S, S, C, K, G, seed=21, mask_div=mdiv)
    cfg = dict(synth.LOSS_CONFIG, seg_w=seg_w, scale_w=[5.0, 2.0, 1.0])
    anc = [synth.anchors_tensor(s).numpy() for s in synth.SCALES]
    ref_loss, ref_m, ref_g, ref_gp = SO.segmentation_loss([p.numpy() for p in preds], t.numpy(), protos.numpy(), masks.numpy(),
                                                          anc, cfg, C, K, with_grad=True)
    loss, metrics, grads, gpr = _seg_case(ops, preds, protos, t, masks, C, K, cfg)
    assert_close(float(loss), ref_loss, rtol=1e-5, atol=0, what="loss")
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for a, b in zip(grads, ref_g):
        assert_close(a, b, rtol=1e-4, atol=1e-8, what="grad preds")
    assert_close(gpr, ref_gp, rtol=1e-4, atol=1e-9, what="grad protos")


def test_segmentation_loss_twice_and_scaled(ops):
    """Two forwards before their backwards (own workspaces), and an upstream gradient other than 1."""
    from tests.util import seg_loss_case
    g = golden("segloss_sq64")
    preds, protos, t, masks, C, K = seg_loss_case("segloss_sq64", g)
    cfg = dict(synth.LOSS_CONFIG, seg_w=1.0)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    gp = [dev(p).requires_grad_(True) for p in preds]
    pr = dev(protos).requires_grad_(True)
    l1, _ = ops.segmentation_loss(gp, dev(t), pr, dev(masks), anc, cfg, C, K, with_metrics=False)
    l2, _ = ops.segmentation_loss([p * 1.0 for p in gp], dev(t), pr * 1.0, dev(masks), anc, cfg, C, K, with_metrics=False)
    (l1 * 0.5 + l2 * 1.5).backward()
    for sc, p in zip(synth.SCALES, gp):
        assert_close(p.grad.cpu().numpy(), 2.0 * g["grad_" + sc], rtol=1e-4, atol=2e-7, what="grad " + sc)
    assert_close(pr.grad.cpu().numpy(), 2.0 * g["grad_protos"], rtol=1e-4, atol=2e-8, what="grad protos")


def test_segmentation_loss_without_targets(ops):
    """No target at all: the mask term is zero (the reference's loop over batch_idx.unique() is empty), the protos get a
    zero gradient, the detection terms are those of the detection loss."""
    from oracle import seg_oracle as SO
    B, S, C, K = 2, 64, 5, 8
    preds, protos, t, masks = synth.seg_inputs(B, S, S, C, K, 0, seed=4, fixed=True)
    assert t.shape == (0, 6)
    cfg = dict(synth.LOSS_CONFIG, seg_w=1.0)
    anc = [synth.anchors_tensor(s).numpy() for s in synth.SCALES]
    ref_loss, ref_m, ref_g, _ = SO.segmentation_loss([p.numpy() for p in preds], t.numpy(), protos.numpy(), masks.numpy(), anc, cfg,
                                                     C, K, with_grad=True)
    loss, metrics, grads, gpr = _seg_case(ops, preds, protos, t, masks, C, K, cfg)
    assert_close(float(loss), ref_loss, rtol=1e-5, atol=0, what="loss")
    assert metrics["seg_loss"] == 0.0 and metrics["dice_score"] == 0.0
    for a, b in zip(grads, ref_g):
        assert_close(a, b, rtol=1e-4, atol=1e-8, what="grad preds")
    assert not gpr.any()


@pytest.mark.parametrize("name", ["segmask_T128", "segmask_T128_odd"])
def test_seg_masks_golden(ops, name):
    """inference_seg.post_process_preds lines 62-117 end to end on the device: rows by ops.post_process, the
    coefficients of the kept rows by ops.extra_columns, the masks by ops.seg_masks -- against the boolean masks the
    UNMODIFIED function hands to its drawing code (pixels may differ only where the interpolated value is within 2e-5 of 0.5)."""
    from oracle import seg_oracle as SO
    from tests.util import assert_masks_match, seg_mask_case
    g = golden(name)
    raws, (B, H, W, C, og), protos, isz, ref_masks = seg_mask_case(g)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    base = torch.cat([ops.decode_scale(dev(r), a, (H, W), True, og).reshape(B, -1, C + 5) for r, a in zip(raws, anc)], 1)
    extra = seg_extra_columns(B, base.shape[1], 4, int(g["extra_seed"]))
    preds = torch.cat([base, dev(extra)], dim=-1).contiguous()
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    det = ops.post_process(preds, (H, W), C, float(g["iou"]), float(g["thr"]), allow, tracked, order="image")
    coefs = ops.extra_columns(preds, det, C)[:, :4].contiguous()
    masks = ops.seg_masks(coefs, det.counts, dev(protos), isz)
    assert masks.dtype == torch.bool and tuple(masks.shape) == (det.pred_boxes.shape[0], isz[0], isz[1])
    counts = g["per_image_counts"]
    ref_img = np.repeat(np.arange(len(counts)), counts)
    rows = det.pred_boxes.cpu().numpy()
    got_img = np.unique(det.sample_idxs.cpu().numpy(), return_inverse=True)[1]
    po, pr = rows_order(rows, got_img), rows_order(g["per_image"], ref_img)
    assert_close(rows[po], g["per_image"][pr], rtol=1e-5, atol=2e-5 * max(H, W), what="rows vs reference")
    _, vals = SO.seg_masks(coefs.cpu().numpy(), det.counts.numpy(), protos.numpy(), isz[0], isz[1])
    nd = assert_masks_match(masks.cpu().numpy()[po], ref_masks[pr], vals[po])
    print("%s: %d masks of %dx%d, %d pixels differ from the reference (all on the threshold)" % (name, masks.shape[0], isz[0], isz[1], nd))


def test_seg_masks_vs_oracle_and_torch(ops):
    """32 coefficients, 160x160 protos to 640x640, images without rows: against the oracle and against the same three torch
    calls on the device (what inference_seg.py runs per image)."""
    from oracle import seg_oracle as SO
    from tests.util import assert_masks_match
    B, K, Hp, Wp, H, W = 4, 32, 160, 160, 640, 640
    g = torch.Generator().manual_seed(5)
    counts = torch.tensor([7, 0, 19, 3])
    coefs = torch.tanh(torch.randn(int(counts.sum()), K, generator=g))
    protos = torch.randn(B, K, Hp, Wp, generator=g)
    masks = ops.seg_masks(dev(coefs), counts, dev(protos), (H, W)).cpu().numpy()
    ref, vals = SO.seg_masks(coefs.numpy(), counts.numpy(), protos.numpy(), H, W)
    nd = assert_masks_match(masks, ref, vals)
    r, outs = 0, []
    for i, c in enumerate(counts.tolist()):
        if c:
            m = (dev(coefs[r:r + c]) @ dev(protos[i]).reshape(K, -1)).reshape(-1, Hp, Wp).sigmoid()
            m = torch.nn.functional.interpolate(m.unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False)
            outs.append(torch.gt(m, 0.5).squeeze(0).cpu().numpy())
        r += c
    nt = assert_masks_match(masks, np.concatenate(outs, 0), vals, what="masks vs torch-CUDA", band=5e-5)
    print("seg_masks 29 x 640x640: %d pixels differ from the oracle, %d from torch-CUDA (all on the threshold)" % (nd, nt))
    assert ops.seg_masks(dev(coefs[:0]), torch.zeros(B, dtype=torch.int64), dev(protos), (H, W)).shape == (0, H, W)


@pytest.mark.parametrize("K,Hp,Wp,H,W,counts", [
    (5, 7, 9, 21, 30, [3, 0, 2]),            # W % 4 != 0: one pixel per thread; K not a multiple of four
    (64, 40, 40, 20, 20, [4, 1]),            # masks made smaller: source lines are skipped; the largest K
    (8, 100, 300, 110, 304, [2, 5]),         # a block's source lines exceed the shared-memory window: read from global memory
    (16, 33, 17, 257, 131, [1]),             # odd everything, enlargement by non-integer factors, a single row
    (3, 2, 2, 9, 8, [0, 0, 6, 0]),           # tiny prototypes (every output line interpolates the same two source lines), empty images
])
def test_seg_masks_shapes(ops, K, Hp, Wp, H, W, counts):
    """bg_seg_masks on the code paths the goldens do not reach: against the numpy restatement of inference_seg.py:115-117
    (a pixel may differ only where the interpolated value is within 2e-5 of 0.5)."""
    from oracle import seg_oracle as SO
    from tests.util import assert_masks_match
    g = torch.Generator().manual_seed(K * 1000 + H)
    counts = torch.tensor(counts)
    B = counts.numel()
    coefs = torch.tanh(torch.randn(int(counts.sum()), K, generator=g))
    protos = torch.randn(B, K, Hp, Wp, generator=g)
    masks = ops.seg_masks(dev(coefs), counts, dev(protos), (H, W))
    assert masks.dtype == torch.bool and tuple(masks.shape) == (int(counts.sum()), H, W)
    ref, vals = SO.seg_masks(coefs.numpy(), counts.numpy(), protos.numpy(), H, W)
    nd = assert_masks_match(masks.cpu().numpy(), ref, vals)
    frac = float(ref.mean())
    assert 0.05 < frac < 0.95                                      # (the case is not trivially all-false / all-true)
    print("seg_masks K=%d %dx%d -> %dx%d rows %s: %d pixels differ (all on the threshold)" % (K, Hp, Wp, H, W, counts.tolist(), nd))
