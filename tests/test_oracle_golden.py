"""CPU suite: the oracle (oracle/boxgeom_oracle.c) against the committed golden fixtures, which are
outputs of the unmodified reference / torchvision-CPU (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from vision_conglomerate_b200 import synth
from tests.util import (ASSIGN_VARIANTS, assign_variant_case, assert_close, canon, digest, golden, rows_canon, rows_order,
                        seg_extra_columns)

DEC = ["dec_sq64", "dec_rect_rescale", "dec_rect_norescale", "dec_T128"]


def _decode_inputs(g):
    B, H, W, C, seed, og0, og1 = (int(v) for v in g["params"])
    raws = synth.raw_head_outputs(B, H, W, C, str(g["dist"]), seed)
    assert digest(*raws) == str(g["in_digest"]), "synthetic generator drifted from the fixture"
    og = None if og0 < 0 else (og0, og1)
    return raws, B, H, W, C, og


@pytest.mark.parametrize("name", DEC)
def test_decode(name):
    g = golden(name)
    raws, B, H, W, C, og = _decode_inputs(g)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = O.decode_inference(raws, anc, H, W, og)
    assert_close(preds[..., C + 1:], g["boxes"], rtol=1e-5, atol=1e-5, what="decoded boxes")
    assert digest(np.ascontiguousarray(preds[..., :C + 1])) == str(g["logits_digest"])  # logits pass through bit-exact
    tr = O.decode_scale(raws[0], anc[0], H, W, inference=False)
    assert_close(tr[..., C + 1:], g["train_sm_boxes"], rtol=1e-5, atol=1e-6, what="training decode")


@pytest.mark.parametrize("name", ["post_sq64", "post_sq64_lowthr", "post_T128_tracked", "post_T128"])
def test_post_process(name):
    g = golden(name)
    gd = golden(str(g["decode_case"]))
    raws, B, H, W, C, og = _decode_inputs(gd)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = O.decode_inference(raws, anc, H, W, og)
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    # NMS stage alone, fed the reference's own fp32 boxes: bit-exact keep list
    N = preds.shape[1]
    score, cls, xyxy = O.score_xyxy(preds, allow or 0.0)
    assert_close(xyxy, g["xyxy"], rtol=1e-5, atol=1e-4, what="xyxy")
    sample = np.repeat(np.arange(B, dtype=np.int64), N)
    keep_ref = g["nms_keep"]
    keep = O.batched_nms(g["xyxy"], score, sample, float(g["iou"]))
    # scores come from our sigmoid; identical ordering is expected unless two scores sit within an ulp
    assert np.array_equal(np.sort(keep), np.sort(keep_ref))
    out = O.post_process(preds, float(g["iou"]), float(g["thr"]), allow, tracked)
    counts = g["per_image_counts"]
    assert out["pred_boxes"].shape[0] == counts.sum()
    # per image, score-descending: compare against what the reference handed to its drawing code
    ref_rows = g["per_image"]
    got = out["pred_boxes"]
    ref_img = np.repeat(np.arange(len(counts)), counts)
    got_img = np.unique(out["sample_idxs"], return_inverse=True)[1] if len(got) else out["sample_idxs"]
    got, ref_rows = rows_canon(got, got_img), rows_canon(ref_rows, ref_img)
    assert_close(got, ref_rows, rtol=1e-5, atol=1e-4, what="pred_boxes")
    assert np.array_equal(got[:, 1], ref_rows[:, 1])  # class ids exact


@pytest.mark.parametrize("name", ["segpost_T128", "segpost_T128_tracked"])
def test_seg_post_process(name):
    """inference_seg.post_process_preds: the box geometry is the detection one on rows that also carry
    mask coefficients; the reference's drawn masks (unit protos: mask = coef > 0) pin the gather of those columns."""
    g = golden(name)
    gd = golden(str(g["decode_case"]))
    raws, B, H, W, C, og = _decode_inputs(gd)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = O.decode_inference(raws, anc, H, W, og)
    extra = seg_extra_columns(B, preds.shape[1], 4, int(g["extra_seed"])).numpy()
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    out = O.post_process(preds, float(g["iou"]), float(g["thr"]), allow, tracked)
    counts = g["per_image_counts"]
    got, ref_rows = out["pred_boxes"], g["per_image"]
    assert got.shape[0] == counts.sum()
    ref_img = np.repeat(np.arange(len(counts)), counts)
    got_img = np.unique(out["sample_idxs"], return_inverse=True)[1] if len(got) else out["sample_idxs"]
    po, pr = rows_order(got, got_img), rows_order(ref_rows, ref_img)
    assert_close(got[po], ref_rows[pr], rtol=1e-5, atol=1e-4, what="pred_boxes")
    coefs = extra.reshape(-1, 4)[out["keep"]]
    assert np.array_equal(coefs[po] > 0, g["masks"][pr].astype(bool))


@pytest.mark.parametrize("case", ["rand", "ties", "dense1", "hand_thr05", "hand_thr05m", "hand_thr0"])
def test_nms(case):
    g = golden("nms")
    b, s, i, thr = g[case + "_boxes"], g[case + "_scores"], g[case + "_idxs"], float(g[case + "_thr"])
    keep = O.batched_nms(b, s, i, thr)
    assert np.array_equal(keep, canon(g[case + "_keep"], s))
    assert np.all(np.diff(s[keep]) <= 0)


def test_nms_live_torchvision():
    import torchvision
    b, s, i = synth.nms_boxes(5000, 7, seed=9)
    ref = torchvision.ops.boxes._batched_nms_vanilla(b, s, i, 0.45).numpy()
    assert np.array_equal(O.batched_nms(b, s, i, 0.45), canon(ref, s.numpy()))
    ref1 = torchvision.ops.nms(b, s, 0.3).numpy()
    assert np.array_equal(O.nms(b, s, 0.3), ref1)


def _assign_cases():
    g = golden("assign")
    keys = sorted(k[:-4] for k in g.files if k.endswith("_idx"))
    return keys


@pytest.mark.parametrize("key", _assign_cases())
def test_assign(key):
    g = golden("assign")
    name, fm, sc = key.rsplit("_", 2)
    ny, nx = (int(v) for v in fm.split("x"))
    t = {"c1": lambda: synth.targets(2, 20, 80, 0, fixed=False), "b8g100": lambda: synth.targets(8, 100, 80, 0),
         "adv": lambda: synth.adversarial_targets(2, 80), "empty": lambda: torch.zeros(0, 6)}[name]()
    assert digest(t) == str(g[name + "_in_digest"])
    idx, cls, anc, box = O.build_target_by_scale(t, (ny, nx), synth.anchors_tensor(sc), 4.0, 0.5)
    assert np.array_equal(np.stack(idx, 0), g[key + "_idx"])          # indices: bit- and order-exact
    assert np.array_equal(cls, g[key + "_cls"])
    assert np.array_equal(anc, g[key + "_anc"])                       # fp32 outputs are bit-reproducible too
    assert np.array_equal(box, g[key + "_box"])


@pytest.mark.parametrize("name", ASSIGN_VARIANTS)
def test_assign_variants(name):
    """Segmentation (overlap_masks) and keypoint-column variants against the reference's outputs."""
    g = golden("assign_variants")
    t, overlap, bs = assign_variant_case(name)
    for (ny, nx), sc in zip(((16, 16), (8, 8), (4, 4)), synth.SCALES):
        idx, cls, anc, box, tm, kp = O.build_target_by_scale_ex(t, (ny, nx), synth.anchors_tensor(sc), 4.0, 0.5, overlap, bs)
        k = f"{name}_{sc}"
        assert np.array_equal(np.stack(idx, 0), g[k + "_idx"]) and np.array_equal(cls, g[k + "_cls"])
        assert np.array_equal(anc, g[k + "_anc"]) and np.array_equal(box, g[k + "_box"])
        assert (tm is None) == (k + "_tmask" not in g.files) and (kp is None) == (k + "_kpts" not in g.files)
        if tm is not None:
            assert np.array_equal(tm, g[k + "_tmask"])
        if kp is not None:
            assert np.array_equal(kp, g[k + "_kpts"])


def test_ciou():
    g = golden("ciou")
    c, grad = O.compute_ciou(g["p"], g["t"], with_grad=True)
    assert_close(c, g["ciou"], rtol=1e-5, atol=1e-6, what="ciou")
    assert_close(grad * g["w"][:, None], g["grad"], rtol=1e-4, atol=1e-5, what="ciou grad")


@pytest.mark.parametrize("name", ["loss_sq64", "loss_collide", "loss_c3_rect", "loss_empty", "loss_c1_640",
                                  "lossraw_sq64", "lossraw_collide", "lossraw_empty", "lossraw_c1_640"])
def test_loss(name):
    """loss_*: the loss on decoded tensors.  lossraw_*: the same tensors as the head's logits, decoded by the
    unmodified _get_scale_pred(inference=False) first, gradients back to the logits (the fused raw form)."""
    g = golden(name)
    B, H, W, C, G, fixed, ts, ps = (int(v) for v in g["params"])
    t = synth.targets(B, G, C, ts, bool(fixed)) if G > 0 else torch.zeros(0, 6)
    preds = synth.train_preds(B, H, W, C, ps)
    assert digest(t, *preds) == str(g["in_digest"])
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    form = "raw" if name.startswith("lossraw_") else "decoded"
    loss, metrics, grads, Ms = O.detection_loss(preds, t, anc, synth.LOSS_CONFIG, with_grad=True, input_form=form)
    if form == "raw":  # the split form is the same computation on the three column groups
        tri = [(p[..., 0].contiguous(), p[..., 1:1 + C].contiguous(), p[..., 1 + C:].contiguous()) for p in preds]
        loss_s, _, grads_s, _ = O.detection_loss(tri, t, anc, synth.LOSS_CONFIG, with_grad=True, input_form="split")
        assert loss_s == loss
        for gr, (gc, gk, gb) in zip(grads, grads_s):
            assert np.array_equal(gr[..., 0], gc) and np.array_equal(gr[..., 1:1 + C], gk) and np.array_equal(gr[..., 1 + C:], gb)
    assert_close(loss, float(g["loss"]), rtol=1e-5, atol=0, what="loss")
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


def test_ratio_metrics():
    g = golden("ratio")
    s = O.ratio_metrics(g["anchors"], g["wh"], 4.0)
    assert_close(s[0], float(g["score"]), rtol=1e-5)
    assert_close(np.array(s), g["extras"], rtol=1e-5)
    assert_close(np.array(O.ratio_metrics(g["anchors"], g["wh"] * 3.0, 2.0)), g["extras_x3_t2"], rtol=1e-5)


@pytest.mark.parametrize("name", ["segloss_sq64", "segloss_rect", "segloss_128", "segloss_empty_img"])
def test_segmentation_loss(name):
    """the numpy restatement of SegmentationLoss.forward (oracle/seg_oracle.py: detection terms from the C
    oracle + the mask term) against the unmodified reference's loss, metrics and gradients with respect to the three
    prediction tensors and the protos (modules/segmentation_loss.py:26-231, overlap_masks=True)."""
    from oracle import seg_oracle as SO
    from tests.util import seg_loss_case
    g = golden(name)
    preds, protos, t, masks, C, K = seg_loss_case(name, g)
    assert digest(t, protos, masks, *preds) == str(g["in_digest"])
    anc = [synth.anchors_tensor(s).numpy() for s in synth.SCALES]
    cfg = dict(synth.LOSS_CONFIG, seg_w=1.0)
    loss, metrics, grads, gp = SO.segmentation_loss([p.numpy() for p in preds], t.numpy(), protos.numpy(), masks.numpy(), anc, cfg,
                                                    C, K, with_grad=True)
    assert_close(loss, float(g["loss"]), rtol=1e-5, atol=0, what="loss")
    ref_m = dict(zip((str(k) for k in g["metric_keys"]), g["metric_vals"]))
    for k, v in ref_m.items():
        assert_close(metrics[k], v, rtol=2e-5, atol=1e-7, what=k)
    for sc, gr in zip(synth.SCALES, grads):
        assert_close(gr, g["grad_" + sc], rtol=1e-4, atol=1e-7, what="grad " + sc)
    assert_close(gp, g["grad_protos"], rtol=1e-4, atol=1e-8, what="grad protos")


@pytest.mark.parametrize("name", ["segmask_T128", "segmask_T128_odd"])
def test_seg_masks(name):
    """inference_seg.post_process_preds lines 115-117: sigmoid(coefs @ protos) -> bilinear resize -> > 0.5, the
    numpy restatement against the boolean masks the unmodified function hands to its drawing code."""
    from oracle import seg_oracle as SO
    from tests.util import assert_masks_match, seg_mask_case
    g = golden(name)
    raws, (B, H, W, C, og), protos, isz, ref_masks = seg_mask_case(g)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    preds = O.decode_inference(raws, anc, H, W, og)
    extra = seg_extra_columns(B, preds.shape[1], 4, int(g["extra_seed"])).numpy()
    allow = None if int(g["allow"]) < 0 else int(g["allow"])
    tracked = [int(v) for v in g["tracked"]] or None
    out = O.post_process(preds, float(g["iou"]), float(g["thr"]), allow, tracked)
    counts = g["per_image_counts"]
    got, ref_rows = out["pred_boxes"], g["per_image"]
    assert got.shape[0] == counts.sum() == ref_masks.shape[0]
    ref_img = np.repeat(np.arange(len(counts)), counts)
    got_img = np.unique(out["sample_idxs"], return_inverse=True)[1] if len(got) else out["sample_idxs"]
    po, pr = rows_order(got, got_img), rows_order(ref_rows, ref_img)
    # image-major rows for the restatement, then back to the reference's order
    img_full = out["sample_idxs"][po]
    coefs = extra.reshape(-1, 4)[out["keep"]][po]
    per_img = np.bincount(img_full, minlength=B)
    m, vals = SO.seg_masks(coefs, per_img, protos.numpy(), isz[0], isz[1])
    nd = assert_masks_match(m, ref_masks[pr], vals)
    print("%s: %d rows, %d of %d pixels differ (all within 2e-5 of the threshold)" % (name, m.shape[0], nd, m.size))
