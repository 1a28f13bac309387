"""GPU suite on the REAL reference (-m gpu): ``dropin.install()`` on the unmodified ``DetectionNet`` /
``DetectionLoss`` / ``DetectionDataset`` / ``inference_det`` of ches-001/vision-conglomerate and patched against
unpatched results on the same B200 (the unpatched run is the reference's own torch-CUDA + torchvision-CUDA path).

The reference travels to the GPU box as ``baseline/_ref/`` (an untouched copy made by ``__graft_entry__.build()`` in
the build container; git-ignored).  Tests skip when it is absent."""
import numpy as np
import pytest
import torch

from oracle import ref_harness
from oracle.parity import explain_keep_mismatches, summarize
from vision_conglomerate_b200 import synth
from tests.util import assert_close, rows_canon

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")]


@pytest.fixture(scope="module")
def ns():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = ref_harness.load()
    n.inference_det.device = "cuda"
    return n


def _model(ns, C=80):
    torch.manual_seed(42)
    m = ns.DetectionNet(3, C, ref_harness.model_config(), synth.ANCHORS, num_keypoints=0)
    return m.cuda()


def _param_grads(model):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def test_train_step_patched_equals_unpatched(ns):
    """One training step of the real DetectionNet + DetectionLoss (pipeline/detection_trainer.py:178-184): model forward,
    loss, backward.  Patched: the head logits reach the fused loss through LazyDecoded stand-ins (no decoded tensor is
    ever built).  Loss, metrics and every parameter gradient must match the unpatched torch-CUDA run."""
    from vision_conglomerate_b200 import dropin, ops, lazy
    B, S, C = 4, 256, 80
    model = _model(ns, C).train()
    loss_mod = ns.DetectionLoss(model, **synth.LOSS_CONFIG)
    g = torch.Generator().manual_seed(3)
    imgs = torch.rand(B, 3, S, S, generator=g).cuda()
    # one target per image, far apart: no two matches share a cell, so the reference's duplicate-index scatter on CUDA
    # (whose winner is not deterministic there) cannot blur the comparison
    t = torch.tensor([[b, (7 * b) % C, 0.2 + 0.15 * b, 0.3 + 0.1 * b, 0.12, 0.3] for b in range(B)], dtype=torch.float32).cuda()

    def step():
        model.zero_grad(set_to_none=True)
        sm, md, lg = model(imgs)
        loss, metrics = loss_mod((sm, md, lg), t)
        loss.backward()
        return float(loss), metrics, _param_grads(model), (sm, md, lg)

    loss_u, met_u, grads_u, _ = step()
    calls = []
    real = ops.detection_loss

    def spy(*a, **k):
        calls.append(k.get("input_form", "decoded"))
        return real(*a, **k)

    dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet, make_anchors=ns.make_anchors)
    ops.detection_loss = spy
    try:
        loss_p, met_p, grads_p, preds_p = step()
        assert calls == ["raw"], calls                                   # the fused path ran, from the logits
        assert all(isinstance(p, lazy.LazyDecoded) and p.pending for p in preds_p)
        # a different consumer of the predictions sees the reference's decoded values (and the stand-in materialises)
        model.zero_grad(set_to_none=True)
        sm, md, lg = model(imgs)
        dec = sm[..., C + 1:].detach()
        assert not sm.pending
        dropin.uninstall()
        sm_u = model(imgs)[0]
        assert_close(dec.cpu().numpy(), sm_u[..., C + 1:].detach().cpu().numpy(), rtol=1e-5, atol=1e-6, what="materialised decode")
        # and the un-fused patched route (real decoded tensors through the differentiable CUDA decode) agrees too
        dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet, fuse_train_decode=False)
        calls.clear()
        loss_d, met_d, grads_d, _ = step()
        assert calls == ["decoded"]
        dropin.uninstall()
        # EffiDecHead's final torch.cat deferred as well -- the loss reads the three conv outputs in place
        dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet, EffiDecHead=ns.EffiDecHead)
        calls.clear()
        loss_s, met_s, grads_s, preds_s = step()
        assert calls == ["split"], calls
        assert all(isinstance(p, lazy.LazyRows) and p.pending and len(p.parts) == 3 for p in preds_s)
        with torch.no_grad():                                             # inference through the patched head still works
            model.eval()
            inf_s = model(imgs, inference=True)
            dropin.uninstall()
            inf_u = model(imgs, inference=True)
            model.train()
        assert_close(inf_s.cpu().numpy(), inf_u.cpu().numpy(), rtol=1e-5, atol=2e-5 * S, what="inference through the deferred head")
    finally:
        ops.detection_loss = real
        dropin.uninstall()
    print("real train step: loss unpatched %.7f, fused %.7f, decoded-route %.7f, split-head %.7f" % (loss_u, loss_p, loss_d, loss_s))
    for lp, mp, gp in ((loss_p, met_p, grads_p), (loss_d, met_d, grads_d), (loss_s, met_s, grads_s)):
        assert_close(lp, loss_u, rtol=1e-5, atol=0, what="loss")
        assert set(mp) == set(met_u)
        for k in met_u:
            assert_close(mp[k], met_u[k], rtol=2e-5, atol=1e-7, what=k)
        assert set(gp) == set(grads_u)
        # all parameter gradients as one vector, and each sizeable one on its own (conv biases in front of a BatchNorm
        # have a mathematically zero gradient: what either run holds there is rounding noise, so parameters whose
        # gradient norm is below 1e-3 of the largest are covered by the global figure only)
        gmax = max(float(g.double().norm()) for g in grads_u.values())
        tot_err = sum(float((gp[n].double() - grads_u[n].double()).pow(2).sum()) for n in grads_u) ** 0.5
        tot_den = sum(float(grads_u[n].double().pow(2).sum()) for n in grads_u) ** 0.5
        worst, worst_name = 0.0, ""
        for n in grads_u:
            den = float(grads_u[n].double().norm())
            if den < 1e-3 * gmax:
                continue
            rel = float((gp[n].double() - grads_u[n].double()).norm()) / den
            if rel > worst:
                worst, worst_name = rel, n
        print("  %d parameter gradients: global relative L2 error %.2e, worst sizeable parameter %.2e (%s)"
              % (len(gp), tot_err / tot_den, worst, worst_name))
        assert tot_err / tot_den < 1e-4 and worst < 5e-4, worst_name


def test_split_head_is_zero_copy_under_channels_last(ns):
    """With the model in channels-last memory format the head's `permute(0,2,3,1).reshape(...)` views of its three conv
    outputs (modules/common.py:912-918) are contiguous: the fused loss reads them where cuDNN wrote them (no copy in
    `_req`) -- and loss / gradients still equal the unpatched run of the same channels-last model."""
    from vision_conglomerate_b200 import dropin, ops
    B, S, C = 2, 256, 80
    model = _model(ns, C).train().to(memory_format=torch.channels_last)
    loss_mod = ns.DetectionLoss(model, **synth.LOSS_CONFIG)
    g = torch.Generator().manual_seed(9)
    imgs = torch.rand(B, 3, S, S, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    t = torch.tensor([[b, 3 + b, 0.3 + 0.2 * b, 0.4, 0.15, 0.35] for b in range(B)], dtype=torch.float32).cuda()

    def step():
        model.zero_grad(set_to_none=True)
        preds = model(imgs)
        loss, _ = loss_mod(preds, t)
        loss.backward()
        return float(loss), _param_grads(model)

    loss_u, grads_u = step()
    seen = []
    real = ops.detection_loss

    def spy(preds3, *a, **k):
        if k.get("input_form") == "split":
            seen.append(all(x.is_contiguous() for tri in preds3 for x in tri))
        return real(preds3, *a, **k)

    dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet, EffiDecHead=ns.EffiDecHead)
    ops.detection_loss = spy
    try:
        loss_p, grads_p = step()
    finally:
        ops.detection_loss = real
        dropin.uninstall()
    assert seen == [True], seen          # the pieces reached the loss contiguous: nothing was copied
    assert_close(loss_p, loss_u, rtol=1e-5, atol=0, what="loss")
    err = sum(float((grads_p[n].double() - grads_u[n].double()).pow(2).sum()) for n in grads_u) ** 0.5
    den = sum(float(grads_u[n].double().pow(2).sum()) for n in grads_u) ** 0.5
    print("channels-last split head: loss %.7f vs %.7f, global relative L2 error of the parameter gradients %.2e" % (loss_p, loss_u, err / den))
    assert err / den < 1e-4


def test_inference_patched_equals_unpatched_and_fused(ns):
    """inference_det.py's own flow (model(x, inference=True, og_size) -> post_process_preds) on the real classes:
    unpatched (ATen decode + torchvision-CUDA batched_nms), patched with install() only (zero edits: CUDA decode,
    CUDA _bbox_to_size, bg_batched_nms over all candidates), and the fused ops.detect from the head outputs."""
    import torchvision
    from vision_conglomerate_b200 import dropin, ops
    B, S, C = 3, 256, 80
    model = _model(ns, C).eval()
    g = torch.Generator().manual_seed(5)
    imgs = torch.rand(B, 3, S, S, generator=g).cuda()
    og = (300, 400)
    iou, thr, allow = 0.5, 0.2, 4
    heads = []
    hooks = [h.register_forward_hook(lambda m, i, o: heads.append(o.detach())) for h in model.head]
    with torch.no_grad():
        preds_u = model(imgs, inference=True, og_size=og)
    for h in hooks:
        h.remove()
    cap_u = ref_harness.ref_post_process(preds_u, C, iou, thr, allow, None)
    dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet)
    try:
        with torch.no_grad():
            preds_p = model(imgs, inference=True, og_size=og)
        cap_p = ref_harness.ref_post_process(preds_p, C, iou, thr, allow, None)
    finally:
        dropin.uninstall()
    assert torchvision.ops.batched_nms.__module__.startswith("torchvision")
    assert_close(preds_p.cpu().numpy(), preds_u.cpu().numpy(), rtol=1e-5, atol=2e-5 * max(og), what="decoded preds")
    N = preds_u.shape[1]
    # same fp32 boxes/scores in, so the keep-lists of the two NMS implementations must be identical sets
    if torch.equal(cap_p["boxes"], cap_u["boxes"]) and torch.equal(cap_p["scores"], cap_u["scores"]):
        assert np.array_equal(np.sort(cap_p["keep"].cpu().numpy()), np.sort(cap_u["keep"].cpu().numpy()))
    rows_u = np.concatenate(cap_u["per_image"]) if cap_u["per_image"] else np.zeros((0, 6), np.float32)
    rows_p = np.concatenate(cap_p["per_image"]) if cap_p["per_image"] else np.zeros((0, 6), np.float32)
    img_u = np.repeat(np.arange(len(cap_u["per_image"])), [len(r) for r in cap_u["per_image"]])
    img_p = np.repeat(np.arange(len(cap_p["per_image"])), [len(r) for r in cap_p["per_image"]])
    par = explain_keep_mismatches(cap_u["scores"].cpu().numpy(), cap_u["boxes"].cpu().numpy(), N,
                                  cap_u["keep"].cpu().numpy()[cap_u["scores"][cap_u["keep"]].cpu().numpy() > thr],
                                  cap_p["keep"].cpu().numpy()[cap_p["scores"][cap_p["keep"]].cpu().numpy() > thr], iou, thr)
    print("real inference, zero-edit vs unpatched: " + summarize(par))
    assert not par["unexplained"]
    if par["mismatches"] == 0:
        assert_close(rows_canon(rows_p, img_p), rows_canon(rows_u, img_u), rtol=1e-5, atol=2e-5 * max(og), what="rows")
    # the fused path from the head outputs
    anchors3 = [model.sm_anchors.data, model.md_anchors.data, model.lg_anchors.data]
    det = ops.detect(heads, anchors3, (S, S), C, og_size=og, iou_threshold=iou, score_threshold=thr, box_allowance=allow,
                     order="global")
    keep_u = cap_u["keep"].cpu().numpy()
    keep_u = keep_u[cap_u["scores"].cpu().numpy()[keep_u] > thr]
    par = explain_keep_mismatches(cap_u["scores"].cpu().numpy(), cap_u["boxes"].cpu().numpy(), N, keep_u,
                                  det.keep_idxs.cpu().numpy(), iou, thr)
    print("real inference, fused ops.detect vs unpatched: " + summarize(par))
    assert not par["unexplained"]
    assert det.keep_idxs.numel() > 0


def test_inference_zero_edit_fused(ns):
    """install(inference_det=module): the unmodified call sequence of inference_det.evaluate_frames --
    preds = model(x, inference=True, og_size=...); post_process_preds(imgs, preds, ...) -- runs on the fused decode+NMS
    kernels with no edit of the reference: the model returns a stand-in, the wrapped post_process_preds runs the fused
    kernels and then the reference's own function on the kept candidates only.  The per-image box arrays handed to the
    reference's drawing code must be those of the unpatched run (torch-CUDA + torchvision-CUDA)."""
    from vision_conglomerate_b200 import _lib, dropin, lazy
    B, S, C = 3, 256, 80
    model = _model(ns, C).eval()
    g = torch.Generator().manual_seed(5)
    imgs = torch.rand(B, 3, S, S, generator=g).cuda()
    for og, iou, thr, allow, tracked in (((300, 400), 0.5, 0.2, 4, None), (None, 0.35, 0.24, None, [1, 4, 7, 16, 17]),
                                         ((256, 400), 0.5, 0.2, 4, None), ((300, 400), 0.5, 0.9, 4, None)):
        with torch.no_grad():
            preds_u = model(imgs, inference=True, og_size=og)
        cap_u = ref_harness.ref_post_process(preds_u, C, iou, thr, allow, tracked)
        dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet, inference_det=ns.inference_det)
        try:
            n0 = _lib.launch_count()
            with torch.no_grad():
                preds_p = model(imgs, inference=True, og_size=og)
            assert isinstance(preds_p, lazy.LazyPreds) and preds_p.pending and tuple(preds_p.shape) == tuple(preds_u.shape)
            assert _lib.launch_count() == n0                       # nothing has been computed yet
            cap_p = ref_harness.ref_post_process(preds_p, C, iou, thr, allow, tracked)
            assert preds_p.pending                                 # ... and the decoded [B,N,85] tensor never was
            assert cap_p["boxes"].shape[0] < preds_u.shape[0] * preds_u.shape[1] // 2   # the reference ran on the kept rows only
        finally:
            dropin.uninstall()
        assert len(cap_p["per_image"]) == len(cap_u["per_image"]), (og, thr)
        nrows = 0
        for a, b in zip(cap_p["per_image"], cap_u["per_image"]):
            assert a.shape == b.shape
            assert_close(rows_canon(a, np.zeros(len(a))), rows_canon(b, np.zeros(len(b))), rtol=1e-5, atol=2e-5 * 400, what="rows")
            assert np.array_equal(np.sort(a[:, 1]), np.sort(b[:, 1]))
            nrows += len(a)
        print("zero-edit fused inference og=%s thr=%.2f tracked=%s: %d rows over %d images, identical to the unpatched run"
              % (og, thr, tracked, nrows, len(cap_p["per_image"])))


def test_ratio_metrics_patched(ns):
    from vision_conglomerate_b200 import dropin
    g = torch.Generator().manual_seed(31)
    wh = 0.01 + 0.5 * torch.rand(5000, 2, generator=g)
    anc = torch.tensor(sum((synth.ANCHORS[s] for s in synth.SCALES), []), dtype=torch.float32)
    ref = ns.make_anchors.ratio_metrics_w_extras(anc, wh, 4.0), ns.make_anchors.ratio_metrics(anc, wh, 4.0)
    dropin.install(make_anchors=ns.make_anchors, torchvision_ops=False)
    try:
        got = ns.make_anchors.ratio_metrics_w_extras(anc.cuda(), wh.cuda(), 4.0), ns.make_anchors.ratio_metrics(anc, wh.cuda(), 4.0)
        host = ns.make_anchors.ratio_metrics(anc, wh, 4.0)      # host tensors keep the reference implementation
    finally:
        dropin.uninstall()
    assert_close(np.array(got[0], dtype=np.float64), np.array([float(v) for v in ref[0]]), rtol=1e-5)
    assert_close(float(got[1]), float(ref[1]), rtol=1e-5)
    assert float(host) == float(ref[1])


def test_segmentation_train_step_patched_equals_unpatched(ns):
    """SURVEY 8 f2 on the REAL classes: one training step of the unmodified SegmentationNet (19 M parameters,
    config/segmentation/config.yaml) + SegmentationLoss (pipeline/segmentation_trainer.py:42-48) -- model forward, loss,
    backward -- with ``dropin.install(SegmentationLoss=...)`` against the unpatched torch-CUDA run: loss, the twelve metrics,
    every parameter gradient."""
    from vision_conglomerate_b200 import dropin, ops
    B, S, C = 4, 256, 5
    torch.manual_seed(42)
    model = ns.SegmentationNet(3, C, ref_harness.model_config("segmentation"), synth.ANCHORS, num_keypoints=0).cuda().train()
    K = model.proto_seg_module.out_channels
    loss_mod = ns.SegmentationLoss(model, overlap_masks=True, **dict(synth.LOSS_CONFIG, seg_w=1.0))
    g = torch.Generator().manual_seed(3)
    imgs = torch.rand(B, 3, S, S, generator=g).cuda()
    # two targets per image, far apart (no duplicate cells: the reference's duplicate-index scatter on CUDA is not deterministic)
    rows = []
    for b in range(B):
        rows += [[b, (2 * b) % C, 0.2 + 0.1 * b, 0.25, 0.14, 0.3], [b, (2 * b + 1) % C, 0.7, 0.6 + 0.05 * b, 0.2, 0.25]]
    t = torch.tensor(rows, dtype=torch.float32).cuda()
    masks = torch.zeros(B, S, S)
    for b in range(B):
        for j, r in enumerate(rows[2 * b: 2 * b + 2]):
            x, y, w, h = r[2:]
            masks[b, int((y - h / 2) * S): int((y + h / 2) * S), int((x - w / 2) * S): int((x + w / 2) * S)] = j + 1
    masks = masks.cuda()

    def step():
        model.zero_grad(set_to_none=True)
        (sm, md, lg), protos = model(imgs)
        loss, metrics = loss_mod((sm, md, lg), t, protos, masks)
        loss.backward()
        return float(loss), metrics, _param_grads(model)

    loss_u, met_u, grads_u = step()
    calls = []
    orig = ops.segmentation_loss
    ops.segmentation_loss = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    dec_calls = []
    orig_dec = ops.decode_train
    ops.decode_train = lambda *a, **k: (dec_calls.append(1), orig_dec(*a, **k))[1]
    # (DetectionNet: SegmentationNet inherits _get_scale_pred -- its training-mode decode, tanh on the coefficients included,
    #  runs as one differentiable CUDA kernel per scale and direction)
    dropin.install(SegmentationLoss=ns.SegmentationLoss, DetectionNet=ns.DetectionNet, torchvision_ops=False)
    try:
        loss_p, met_p, grads_p = step()
        assert len(dec_calls) == 3, dec_calls
    finally:
        dropin.uninstall()
        ops.segmentation_loss = orig
        ops.decode_train = orig_dec
    assert calls, "the CUDA segmentation loss did not run"
    assert K == 32
    assert_close(loss_p, loss_u, rtol=1e-5, atol=0, what="loss")
    assert set(met_p) == set(met_u)
    for k in met_u:
        assert_close(met_p[k], met_u[k], rtol=5e-5, atol=1e-6, what=k)
    assert set(grads_p) == set(grads_u)
    num = sum(float((grads_p[n].double() - grads_u[n].double()).pow(2).sum()) for n in grads_u)
    den = sum(float(grads_u[n].double().pow(2).sum()) for n in grads_u)
    print("segmentation train step: %d parameter gradients, global relative L2 error %.2e" % (len(grads_u), (num / den) ** 0.5))
    assert (num / den) ** 0.5 < 2e-5


def test_segmentation_inference_zero_edit(ns):
    """SURVEY 8 f2, inference side, on the REAL classes: the unmodified call sequence of inference_seg.evaluate_frames --
    ``preds, protos = model(x, inference=True, og_size=...)``; ``post_process_preds(imgs, preds, protos, ...)`` -- with
    ``dropin.install(inference_seg=module)``: NMS by bg_batched_nms, the per-image boolean masks of lines 115-117 by the
    two mask kernels (recognised through lazy.ProtoTrace / lazy.LazyMasks).  Box arrays and masks handed to the
    reference's drawing code against the unpatched torch-CUDA run; a mask pixel may differ only where the unpatched
    run's interpolated value is within 5e-5 of 0.5."""
    import inference_seg
    from vision_conglomerate_b200 import _lib, dropin, lazy, ops
    B, S, C = 3, 128, 5       # (a random-init model keeps hundreds of rows per image: small frames keep the mask arrays small)
    torch.manual_seed(42)
    model = ns.SegmentationNet(3, C, ref_harness.model_config("segmentation"), synth.ANCHORS, num_keypoints=0).cuda().eval()
    imgs = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(5)).cuda()
    for og, iou, thr, tracked in (((150, 200), 0.5, 0.2, None), ((128, 128), 0.35, 0.24, [0, 2, 3])):
        with torch.no_grad():
            preds, protos = model(imgs, inference=True, og_size=og)
        # the model's own inference decode (three scales: _get_scale_pred with tanh on the mask coefficients, _bbox_to_size)
        # on the decode kernels: install(DetectionNet=...) patches the methods SegmentationNet inherits
        dropin.install(DetectionNet=ns.DetectionNet, torchvision_ops=False)
        try:
            n0 = _lib.launch_count()
            with torch.no_grad():
                preds_p, protos_p = model(imgs, inference=True, og_size=og)
            dec_launches = _lib.launch_count() - n0
        finally:
            dropin.uninstall()
        assert dec_launches >= 3 and tuple(preds_p.shape) == tuple(preds.shape) and torch.equal(protos_p, protos)
        assert_close(preds_p.cpu().numpy(), preds.cpu().numpy(), rtol=1e-5, atol=2e-5 * max(og), what="segmentation decode")
        cap_u = ref_harness.ref_seg_post_process(preds, protos, C, iou, thr, 4, tracked, img_size=og, capture_values=True)
        for mode in ("masks", "fused"):
            calls = []
            orig = ops.seg_masks
            ops.seg_masks = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
            # "masks": post_process_preds on the decoded tensor (NMS by bg_batched_nms over all candidates, masks by the mask
            # kernels); "fused": the model hands over a stand-in, the wrapped function runs the fused decode+NMS on the head
            # outputs and the reference's code on the kept rows only
            if mode == "masks":
                dropin.install(inference_seg=inference_seg)
            else:
                dropin.install(DetectionNet=ns.DetectionNet, inference_seg=inference_seg)
            try:
                n0 = _lib.launch_count()
                if mode == "fused":
                    with torch.no_grad():
                        preds_l, protos_l = model(imgs, inference=True, og_size=og)
                    assert isinstance(preds_l, lazy.LazyPreds) and preds_l.pending and tuple(preds_l.shape) == tuple(preds.shape)
                    assert _lib.launch_count() == n0                   # nothing decoded yet
                    cap_p = ref_harness.ref_seg_post_process(preds_l, protos_l, C, iou, thr, 4, tracked, img_size=og)
                    assert preds_l.pending                             # ... and the decoded [B, N, D] tensor never was
                    assert cap_p["boxes"].shape[0] < preds.shape[0] * preds.shape[1]   # the reference ran on the kept rows only
                else:
                    cap_p = ref_harness.ref_seg_post_process(preds, protos, C, iou, thr, 4, tracked, img_size=og)
                launches = _lib.launch_count() - n0
            finally:
                dropin.uninstall()
                ops.seg_masks = orig
            assert len(cap_u["masks"]) > 0 and len(calls) == len(cap_u["masks"]), (len(calls), len(cap_u["masks"]))
            assert len(cap_p["per_image"]) == len(cap_u["per_image"]) and len(cap_p["masks"]) == len(cap_u["masks"])
            if mode == "masks":
                assert np.array_equal(np.sort(cap_p["keep"].cpu().numpy()), np.sort(cap_u["keep"].cpu().numpy()))
            nd = npx = bad = 0
            for a, b, ma, mb, v in zip(cap_p["per_image"], cap_u["per_image"], cap_p["masks"], cap_u["masks"], cap_u["values"]):
                assert a.shape == b.shape and ma.shape == mb.shape == v.shape and ma.dtype == np.bool_
                oa, ob = np.lexsort((a[:, 2], a[:, 1], -a[:, 0])), np.lexsort((b[:, 2], b[:, 1], -b[:, 0]))
                assert_close(a[oa], b[ob], rtol=1e-5, atol=2e-5 * 400, what="rows")
                diff = ma[oa] != mb[ob]
                nd += int(diff.sum())
                npx += diff.size
                bad += int((diff & (np.abs(v[ob].astype(np.float64) - 0.5) >= 5e-5)).sum())
            assert bad == 0, "%d mask pixels differ away from the threshold" % bad
            print("zero-edit segmentation inference (%s) og=%s tracked=%s: %d images, %d masks, %d of %d pixels differ (all on the "
                  "threshold), %d launches of ours" % (mode, og, tracked, len(cap_u["masks"]), sum(len(m) for m in cap_u["masks"]), nd, npx, launches))
