"""Generate tests/golden/*.npz by running the UNMODIFIED reference (and the installed torchvision CPU
NMS it depends on) on seeded synthetic inputs.  Run in the build container only:

    python oracle/make_golden.py

Inputs are regenerated from ``vision_conglomerate_b200.synth`` seeds by the tests; each fixture stores
the generation parameters, a checksum of the inputs and the reference's outputs.
TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from vision_conglomerate_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def digest(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t).tobytes())
    return h.hexdigest()[:16]


def save(name, **kw):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in kw.items()})


DECODE_CASES = {
    # name: (B, H, W, C, dist, seed, og_size)
    "dec_sq64": (2, 64, 64, 80, "N", 11, None),
    "dec_rect_rescale": (2, 96, 64, 3, "N", 12, (120, 100)),
    "dec_rect_norescale": (1, 96, 64, 3, "N", 13, (96, 100)),   # `and` guard at detection.py:76 -> no rescale
    "dec_T128": (3, 128, 128, 80, "TP", 7, None),
}

POST_CASES = {
    # name: (decode case, iou, score_thr, box_allowance, tracked)
    "post_sq64": ("dec_sq64", 0.65, 0.3, 4, None),
    "post_sq64_lowthr": ("dec_sq64", 0.35, 0.001, None, None),
    "post_T128_tracked": ("dec_T128", 0.35, 0.3, 4, (1, 4, 7, 16, 17)),
    "post_T128": ("dec_T128", 0.65, 0.001, 4, None),
}


def gen_decode_post(ns):
    decoded = {}
    for name, (B, H, W, C, dist, seed, og) in DECODE_CASES.items():
        raws = synth.raw_head_outputs(B, H, W, C, dist, seed)
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        preds = ref_harness.ref_decode_inference(raws, anc, H, W, og, C)
        decoded[name] = (preds, C)
        # training-mode decode of the sm scale as well
        m = ns.DecodeOnly(C)
        with torch.no_grad():
            tr = m._get_scale_pred(raws[0].clone(), anc[0], input_shape=(H, W), inference=False)
        save(name, params=np.array([B, H, W, C, seed, -1 if og is None else og[0], -1 if og is None else og[1]]),
             dist=np.array(dist), in_digest=np.array(digest(*raws)),
             boxes=preds[..., C + 1:C + 5].numpy(), logits_digest=np.array(digest(preds[..., :C + 1].contiguous())),
             train_sm_boxes=tr[..., C + 1:C + 5].numpy())
    for name, (dname, iou, thr, allow, tracked) in POST_CASES.items():
        preds, C = decoded[dname]
        cap = ref_harness.ref_post_process(preds, C, iou, thr, allow, tracked)
        B = preds.shape[0]
        # post_process_preds gives the drawing code one array per surviving image, in image order,
        # skipping images left empty (inference_det.py:100-112).  Recover the image id from the NMS capture.
        keep = cap["keep"]
        sc = cap["scores"][keep]
        m = sc > thr
        simg = cap["idxs"][keep][m]
        per = cap["per_image"]
        save(name, decode_case=np.array(dname), iou=np.array(iou), thr=np.array(thr),
             allow=np.array(-1 if allow is None else allow),
             tracked=np.array(tracked if tracked else [], dtype=np.int64),
             nms_keep=keep.numpy(), nms_scores_digest=np.array(digest(cap["scores"])),
             xyxy=cap["boxes"].numpy() if cap["boxes"].numel() < 200000 else cap["boxes"][keep].numpy(),
             scores_kept=sc.numpy(), kept_img=simg.numpy(),
             n_images=np.array(len(per)),
             per_image=np.concatenate(per, 0) if per else np.zeros((0, 6), np.float32),
             per_image_counts=np.array([p.shape[0] for p in per], dtype=np.int64))


SEG_POST_CASES = {
    # name: (decode case, iou, score_thr, box_allowance, tracked, keypoint columns)
    "segpost_T128": ("dec_T128", 0.65, 0.001, 4, None, 0),
    "segpost_T128_tracked": ("dec_T128", 0.35, 0.3, 4, (1, 4, 7, 16, 17), 0),   # (keypoint columns next to masks trip the reference's own assert, inference_seg.py:70)
}
SEG_MASKS = 4


def seg_extra_columns(B, N, n_kp, seed):
    """Mask coefficients (tanh range) and keypoint columns appended to the decoded rows, seeded."""
    g = torch.Generator().manual_seed(seed)
    return torch.tanh(torch.randn(B, N, SEG_MASKS + n_kp, generator=g))


def gen_seg_post(ns):
    """inference_seg.post_process_preds (f2): same box geometry on rows that carry mask coefficients (and keypoints).
    The protos are the 4 unit masks over a 2x2 map and the image is 2x2, so the masks the reference draws are
    `coef > 0` of the kept rows: they pin the row gather of the extra columns."""
    for name, (dname, iou, thr, allow, tracked, n_kp) in SEG_POST_CASES.items():
        B, H, W, C, dist, seed, og = DECODE_CASES[dname]
        raws = synth.raw_head_outputs(B, H, W, C, dist, seed)
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        base = ref_harness.ref_decode_inference(raws, anc, H, W, og, C)
        extra = seg_extra_columns(B, base.shape[1], n_kp, 500 + seed)
        preds = torch.cat([base, extra], dim=-1).contiguous()
        protos = torch.eye(SEG_MASKS).reshape(1, SEG_MASKS, 2, 2).repeat(B, 1, 1, 1).contiguous()
        cap = ref_harness.ref_seg_post_process(preds, protos, C, iou, thr, allow, tracked)
        per, masks, kps = cap["per_image"], cap["masks"], cap["keypoints"]
        save(name, decode_case=np.array(dname), iou=np.array(iou), thr=np.array(thr),
             allow=np.array(-1 if allow is None else allow), n_kp=np.array(n_kp), extra_seed=np.array(500 + seed),
             tracked=np.array(tracked if tracked else [], dtype=np.int64),
             nms_keep=cap["keep"].numpy(),
             per_image=np.concatenate(per, 0) if per else np.zeros((0, 6), np.float32),
             per_image_counts=np.array([p.shape[0] for p in per], dtype=np.int64),
             masks=np.concatenate([m.reshape(m.shape[0], -1) for m in masks], 0) if masks else np.zeros((0, SEG_MASKS), bool),
             keypoints=np.concatenate([k.reshape(-1, 3) for k in kps], 0) if kps else np.zeros((0, 3), np.float32),
             keypoint_rows=np.array([k.reshape(-1, 3).shape[0] for k in kps], dtype=np.int64))


SEG_MASK_CASES = {
    # name: (decode case, iou, score_thr, box_allowance, tracked, protos (Hp, Wp), image (H, W))
    "segmask_T128": ("dec_T128", 0.35, 0.3, 4, None, (32, 32), (128, 128)),          # x4 bilinear
    "segmask_T128_odd": ("dec_T128", 0.35, 0.3, 4, (1, 4, 7, 16, 17), (24, 40), (90, 100)),  # non-integer scales, class filter
}


def gen_seg_masks(ns):
    """inference_seg.post_process_preds lines 115-117 (f2): masks = sigmoid(coefs @ protos) -> bilinear resize to the image
    -> > 0.5, captured from the unmodified function for every surviving image (boolean, bit-packed in the fixture)."""
    for name, (dname, iou, thr, allow, tracked, psz, isz) in SEG_MASK_CASES.items():
        B, H, W, C, dist, seed, og = DECODE_CASES[dname]
        raws = synth.raw_head_outputs(B, H, W, C, dist, seed)
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        base = ref_harness.ref_decode_inference(raws, anc, H, W, og, C)
        extra = seg_extra_columns(B, base.shape[1], 0, 500 + seed)
        preds = torch.cat([base, extra], dim=-1).contiguous()
        protos = torch.randn(B, SEG_MASKS, psz[0], psz[1], generator=torch.Generator().manual_seed(900 + seed)).contiguous()
        cap = ref_harness.ref_seg_post_process(preds, protos, C, iou, thr, allow, tracked, img_size=isz)
        per, masks = cap["per_image"], cap["masks"]
        allm = np.concatenate([m.reshape(m.shape[0], -1) for m in masks], 0) if masks else np.zeros((0, isz[0] * isz[1]), bool)
        save(name, decode_case=np.array(dname), iou=np.array(iou), thr=np.array(thr), allow=np.array(-1 if allow is None else allow),
             extra_seed=np.array(500 + seed), proto_seed=np.array(900 + seed), psz=np.array(psz), isz=np.array(isz),
             tracked=np.array(tracked if tracked else [], dtype=np.int64),
             per_image=np.concatenate(per, 0) if per else np.zeros((0, 6), np.float32),
             per_image_counts=np.array([p.shape[0] for p in per], dtype=np.int64),
             masks_packed=np.packbits(allm.astype(np.uint8), axis=1), mask_rows=np.array(allm.shape[0]))


def gen_nms():
    cases = {}
    b, s, g = synth.nms_boxes(3000, 4, seed=3)
    cases["rand"] = (b, s, g, 0.5)
    b, s, g = synth.nms_boxes(3000, 5, seed=4, ties=True)
    cases["ties"] = (b, s, g, 0.5)
    b, s, g = synth.nms_boxes(2000, 1, seed=5, extent=100.0)
    cases["dense1"] = (b, s, g, 0.65)
    # hand-made: IoU exactly 0.5 against thr 0.5 and 0.5-1e-12; zero-area; negative extent; 3-way tie
    hb = torch.tensor([[0, 0, 2, 1], [1, 0, 3, 1],  # inter 1, union 3 -> 1/3
                       [0, 0, 2, 2], [0, 0, 2, 1],  # inter 2, union 4 -> exactly 0.5
                       [5, 5, 5, 5], [5, 5, 5, 5],  # zero area: 0/0 NaN never suppresses
                       [9, 9, 8, 8], [8, 8, 9, 9],  # negative extent
                       [20, 20, 30, 30], [20, 20, 30, 30], [20, 20, 30, 30]], dtype=torch.float32)
    hs = torch.tensor([0.9, 0.8, 0.7, 0.6, 0.5, 0.5, 0.4, 0.3, 0.2, 0.2, 0.2])
    hg = torch.zeros(11, dtype=torch.int64)
    cases["hand_thr05"] = (hb, hs, hg, 0.5)
    cases["hand_thr05m"] = (hb, hs, hg, 0.5 - 1e-12)
    cases["hand_thr0"] = (hb, hs, hg, 0.0)
    out = {}
    for name, (b, s, g, thr) in cases.items():
        keep = torchvision.ops.boxes._batched_nms_vanilla(b, s, g, thr)
        keep2 = torchvision.ops.batched_nms(b, s, g, thr) if b.numel() > 4000 else keep
        assert sorted(keep.tolist()) == sorted(keep2.tolist())
        out[name + "_boxes"] = b.numpy()
        out[name + "_scores"] = s.numpy()
        out[name + "_idxs"] = g.numpy()
        out[name + "_thr"] = np.array(thr)
        out[name + "_keep"] = keep.numpy()
    save("nms", **out)


ASSIGN_CASES = {
    # name: (targets factory, list of fmap shapes)
    "c1": (lambda: synth.targets(2, 20, 80, 0, fixed=False), [(80, 80), (40, 40), (20, 20)]),
    "b8g100": (lambda: synth.targets(8, 100, 80, 0), [(80, 80), (40, 40), (20, 20)]),
    "adv": (lambda: synth.adversarial_targets(2, 80), [(80, 80), (40, 40), (20, 20), (12, 8)]),
    "empty": (lambda: torch.zeros(0, 6), [(20, 20)]),
}


def gen_assign(ns):
    out = {}
    for name, (mk, fmaps) in ASSIGN_CASES.items():
        t = mk()
        out[name + "_in_digest"] = np.array(digest(t))
        for (ny, nx) in fmaps:
            for sc in synth.SCALES:
                anc = synth.anchors_tensor(sc)
                idx, cls, a, box, _, _ = ns.DetectionDataset.build_target_by_scale(t.clone(), (ny, nx), anc, 4.0, 0.5)
                k = f"{name}_{ny}x{nx}_{sc}"
                out[k + "_idx"] = torch.stack(idx, 0).numpy() if cls.numel() else np.zeros((4, 0), np.int64)
                out[k + "_cls"] = cls.numpy()
                out[k + "_anc"] = a.numpy().reshape(-1, 2)
                out[k + "_box"] = box.numpy().reshape(-1, 4)
    save("assign", **out)


def gen_assign_variants(ns):
    """Segmentation (overlap_masks) and keypoint-column variants of build_target_by_scale."""
    out = {}
    cases = {
        "seg_overlap": (synth.targets(4, 12, 80, 3, fixed=False), True, 4),
        "seg_plain": (synth.targets(4, 12, 80, 3, fixed=False), False, None),
        "kpt": (synth.keypoint_targets(3, 10, 2), None, None),
        "kpt_seg_overlap": (synth.keypoint_targets(3, 10, 2), True, 3),
    }
    for name, (t, overlap, bs) in cases.items():
        for (ny, nx), sc in zip(((16, 16), (8, 8), (4, 4)), synth.SCALES):
            anc = synth.anchors_tensor(sc)
            idx, cls, a, box, tm, kp = ns.DetectionDataset.build_target_by_scale(t.clone(), (ny, nx), anc, 4.0, 0.5, overlap, bs)
            k = f"{name}_{sc}"
            out[k + "_idx"] = torch.stack(idx, 0).numpy() if cls.numel() else np.zeros((4, 0), np.int64)
            out[k + "_cls"] = cls.numpy()
            out[k + "_anc"] = a.numpy().reshape(-1, 2)
            out[k + "_box"] = box.numpy().reshape(-1, 4)
            if tm is not None:
                out[k + "_tmask"] = tm.numpy()
            if kp is not None:
                out[k + "_kpts"] = kp.numpy()
    save("assign_variants", **out)


def gen_ciou(ns):
    g = torch.Generator().manual_seed(21)
    M = 512
    p = torch.cat([torch.rand(M, 2, generator=g) * 1.5 - 0.25, torch.rand(M, 2, generator=g) * 6 + 0.05], 1)
    t = torch.cat([torch.rand(M, 2, generator=g), torch.rand(M, 2, generator=g) * 6 + 0.05], 1)
    p[:8] = t[:8]                      # identical boxes
    p[8:16, 2:] = t[8:16, 2:]          # same size, shifted
    p.requires_grad_(True)
    c = ns.DetectionLoss.compute_ciou(p, t)
    w = torch.linspace(0.5, 1.5, M)
    (c * w).sum().backward()
    save("ciou", p=p.detach().numpy(), t=t.numpy(), ciou=c.detach().numpy(), w=w.numpy(), grad=p.grad.numpy())


LOSS_CASES = {
    # name: (B, H, W, C, G, fixed, targets seed, preds seed)
    "loss_sq64": (2, 64, 64, 80, 6, False, 0, 1),
    "loss_collide": (2, 64, 64, 80, 40, True, 2, 3),     # 40 gt on 8x8/4x4/2x2 maps: heavy duplicate cells
    "loss_c3_rect": (3, 96, 64, 3, 10, True, 4, 5),
    "loss_empty": (2, 64, 64, 80, 0, True, 0, 1),
    "loss_c1_640": (2, 640, 640, 80, 20, False, 0, 1),
}


def gen_loss(ns, raw=False):
    """raw=False: the loss on given (already decoded) tensors.  raw=True: the same seeded tensors taken as the
    HEAD's outputs -- the unmodified ``DetectionNet._get_scale_pred(inference=False)`` (modules/detection.py:98-173)
    decodes them first, exactly as ``DetectionNet.forward`` does, and the gradient is taken back to the logits."""
    torch.set_num_threads(1)  # the reference's duplicate-index scatter is racy with more
    for name, (B, H, W, C, G, fixed, ts, ps) in LOSS_CASES.items():
        if raw and name not in RAW_LOSS_CASES:
            continue
        t = synth.targets(B, G, C, ts, fixed) if G > 0 else torch.zeros(0, 6)
        preds = [p.requires_grad_(True) for p in synth.train_preds(B, H, W, C, ps)]
        loss_mod = ns.DetectionLoss(ns.FakeModel(C, synth.ANCHORS), **synth.LOSS_CONFIG)
        if raw:
            net = ns.DecodeOnly(C)
            dec = tuple(net._get_scale_pred(p, synth.anchors_tensor(sc), input_shape=(H, W), inference=False)
                        for p, sc in zip(preds, synth.SCALES))
            loss, metrics = loss_mod(dec, t.clone())
        else:
            loss, metrics = loss_mod(tuple(preds), t.clone())
        loss.backward()
        kw = dict(params=np.array([B, H, W, C, G, int(fixed), ts, ps]), in_digest=np.array(digest(t, *preds)),
                  loss=np.array(loss.item(), np.float64),
                  metric_keys=np.array(list(metrics.keys())),
                  metric_vals=np.array([float(v) for v in metrics.values()], np.float64))
        big = preds[0].numel() > 400000
        for sc, p in zip(synth.SCALES, preds):
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if big:  # keep the fixture small: objectness plane checksums + sparse box/class grads
                kw["grad_" + sc + "_obj_sum"] = np.array(g[..., 0].double().sum().item())
                kw["grad_" + sc + "_abs_sum"] = np.array(g.double().abs().sum().item())
                nz = (g[..., 1:].abs().sum(-1) > 0).nonzero()
                sel = nz[:: max(1, nz.shape[0] // 64)]
                kw["grad_" + sc + "_rows_idx"] = sel.numpy()
                kw["grad_" + sc + "_rows"] = g[sel[:, 0], sel[:, 1], sel[:, 2], sel[:, 3]].numpy()
            else:
                kw["grad_" + sc] = g.numpy()
        save(("lossraw_" + name[5:]) if raw else name, **kw)


RAW_LOSS_CASES = ("loss_sq64", "loss_collide", "loss_empty", "loss_c1_640")


def gen_ratio(ns):
    g = torch.Generator().manual_seed(31)
    wh = 0.01 + 0.5 * torch.rand(1000, 2, generator=g)
    anc = torch.tensor(sum((synth.ANCHORS[s] for s in synth.SCALES), []), dtype=torch.float32)
    s1 = ns.make_anchors.ratio_metrics(anc, wh, 4.0)
    s2 = ns.make_anchors.ratio_metrics_w_extras(anc, wh, 4.0)
    s3 = ns.make_anchors.ratio_metrics_w_extras(anc, wh * 3.0, 2.0)
    save("ratio", wh=wh.numpy(), anchors=anc.numpy(), score=np.array(s1), extras=np.array(s2), extras_x3_t2=np.array(s3))


SEG_LOSS_CASES = {
    # name: (B, H, W, C, K, G, seed, mask_div, fixed)
    "segloss_sq64": (2, 64, 64, 5, 8, 6, 3, 1, False),        # masks 64x64 -> nearest-resized to the 32x32 protos
    "segloss_rect": (3, 96, 64, 80, 32, 4, 5, 2, False),      # masks already at the protos' size, 32 coefficients
    "segloss_128": (2, 128, 128, 80, 32, 12, 7, 1, True),
    "segloss_empty_img": (3, 64, 64, 5, 8, 3, 9, 1, False),   # (image 1 loses its targets below)
}


def gen_seg_loss(ns):
    """``SegmentationLoss.forward`` + backward (modules/segmentation_loss.py:26-231, overlap_masks=True, BCE) on seeded
    inputs: loss, the metrics dict, and the gradients with respect to the three prediction tensors and the protos."""
    torch.set_num_threads(1)
    cfg = dict(synth.LOSS_CONFIG, seg_w=1.0)
    for name, (B, H, W, C, K, G, seed, mdiv, fixed) in SEG_LOSS_CASES.items():
        preds, protos, t, masks = synth.seg_inputs(B, H, W, C, K, G, seed, mdiv, fixed)
        if name == "segloss_empty_img":
            keep = t[:, 0] != 1
            t = t[keep].contiguous()
        preds = [p.requires_grad_(True) for p in preds]
        protos.requires_grad_(True)
        mod = ns.SegmentationLoss(ns.FakeSegModel(C, synth.ANCHORS, K), overlap_masks=True, **cfg)
        if name == "segloss_empty_img":
            # the reference numbers the overlapped masks by the per-image counts for ids 0..B-1 (detection_dataset.py:151-154)
            pass
        loss, metrics = mod(tuple(preds), t.clone(), protos, masks.clone())
        loss.backward()
        kw = dict(params=np.array([B, H, W, C, K, G, seed, mdiv, int(fixed)]), in_digest=np.array(digest(t, protos, masks, *preds)),
                  loss=np.array(loss.item(), np.float64), metric_keys=np.array(list(metrics.keys())),
                  metric_vals=np.array([float(v) for v in metrics.values()], np.float64),
                  grad_protos=protos.grad.numpy())
        for sc, p in zip(synth.SCALES, preds):
            kw["grad_" + sc] = p.grad.numpy()
        save(name, **kw)


if __name__ == "__main__":
    ns = ref_harness.load()
    gen_decode_post(ns)
    gen_seg_post(ns)
    gen_nms()
    gen_assign(ns)
    gen_assign_variants(ns)
    gen_ciou(ns)
    gen_loss(ns)
    gen_loss(ns, raw=True)
    gen_ratio(ns)
    gen_seg_loss(ns)
    gen_seg_masks(ns)
    print("torch", torch.__version__, "torchvision", torchvision.__version__)
