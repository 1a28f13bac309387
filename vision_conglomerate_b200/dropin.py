"""Drop-in installation under the reference's own Python call sites.

The reference has no plugin registry: ``train_det.py`` / ``inference_det.py`` reach the hot path through
attribute look-ups resolved at call time (``DetectionDataset.build_target_by_scale``,
``DetectionLoss.compute_ciou`` / ``forward``, ``torchvision.ops.batched_nms``,
``DetectionNet._get_scale_pred``).  :func:`install` re-points those attributes at the CUDA operators,
keeping every signature and return contract, so the host scripts run unmodified.

Variants that are out of scope for the CUDA path (keypoint columns, focal loss, the segmentation model's loss)
are delegated to the reference's *own original callable*, which is saved at install time -- never to a
re-implementation of ours.  CPU tensors are refused: there is no CPU fallback.

Training (``train_det.py``): ``_get_scale_pred(inference=False)`` returns a :class:`lazy.LazyDecoded` stand-in and
``DetectionLoss.forward`` feeds the head's logits to the fused loss (decode applied in registers, gradient back to
the logits); nothing else in the reference's step touches the predictions (pipeline/detection_trainer.py:178-184).
Inference (``inference_det.py``): with the module passed to ``install(inference_det=...)`` the whole of
``model(x, inference=True, og_size) -> post_process_preds(...)`` runs on the fused decode+NMS kernels with zero edits
(the model returns a :class:`lazy.LazyPreds` stand-in; ``post_process_preds`` is wrapped: fused kernels first, then
the reference's OWN function on the kept candidates only, so its drawing / tracking / CSV code stays its own).
Without it: decode and ``_bbox_to_size`` per scale on the device, ``batched_nms`` on the segmented engine.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import ops
import threading

from .lazy import HeadTrace, LazyPreds, LazyRows, ProtoTrace, loss_inputs_if_pending

_saved: Dict[str, Any] = {}


def _is_cuda_f32(t) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32


def _need_cuda(t, what: str) -> None:
    if not _is_cuda_f32(t):
        raise RuntimeError(f"{what}: vision_conglomerate_b200 is installed and only handles CUDA fp32 tensors "
                           "(no CPU fallback); call uninstall() to get the reference implementation back")


# ------------------------------------------------------------------------------------------------ B1
def _make_build_target_by_scale(orig):
    def build_target_by_scale(targets, fmap_shape, anchors, anchor_threshold: float = 4.0,
                              edge_threshold: float = 0.5, overlap_masks: Optional[bool] = None,
                              batch_size: Optional[int] = None):
        # every variant (detection, segmentation masks, keypoint columns) runs on the CUDA kernel
        _need_cuda(targets, "build_target_by_scale")
        return ops.build_target_by_scale(targets, fmap_shape, anchors, anchor_threshold, edge_threshold, overlap_masks,
                                         batch_size)
    return build_target_by_scale


# ------------------------------------------------------------------------------------------------ B2
def _make_compute_ciou(orig):
    def compute_ciou(preds_xywh, targets_xywh, e: float = 1e-7):
        if targets_xywh.requires_grad and preds_xywh.shape != targets_xywh.shape:
            return orig(preds_xywh, targets_xywh, e)   # (gradient to broadcast targets: never asked for by the reference)
        _need_cuda(preds_xywh, "compute_ciou")
        return ops.compute_ciou(preds_xywh, targets_xywh.to(preds_xywh.dtype), e)
    return compute_ciou


# ------------------------------------------------------------------------------------------------ B3
def _make_loss_forward(orig):
    def forward(self, preds, targets):
        model = self.model
        out_of_scope = (
            (self.alpha and self.gamma)                       # FocalLoss configured (detection_loss.py:74-76)
            or bool(getattr(model, "num_keypoints", None))    # keypoint branch (:147-173)
            or hasattr(model, "proto_seg_module")             # segmentation model
            or len(preds) != 3
            or (isinstance(targets, torch.Tensor) and targets.dim() == 2 and targets.shape[1] > 6)
            or type(self).loss_fn is not _saved.get("DetectionLoss.loss_fn", type(self).loss_fn)
        )
        if out_of_scope:
            return orig(self, preds, targets)   # (LazyDecoded predictions materialise themselves there)
        for p in preds:
            _need_cuda(p, "DetectionLoss.forward")
        cfg = dict(anchor_t=self.anchor_t, edge_t=self.edge_t, box_w=self.box_w, conf_w=self.conf_w,
                   class_w=self.class_w, label_smoothing=self.label_smoothing, scale_w=self.scale_w,
                   batch_scale_loss=self.batch_scale_loss)
        # anchors are read from the module at call time: they live in the state dict (detection.py:36-38)
        anchors3 = [model.sm_anchors.data, model.md_anchors.data, model.lg_anchors.data]
        targets = targets.to(preds[0].device, torch.float32)
        lazy_in = loss_inputs_if_pending(preds)
        if lazy_in is not None:   # straight from the head: the training-mode decode is fused into the loss
            return ops.detection_loss(lazy_in[1], targets, anchors3, cfg, input_form=lazy_in[0])
        return ops.detection_loss(preds, targets, anchors3, cfg)
    return forward


# ------------------------------------------------------------------------------------------------ f2
def _make_seg_loss_forward(orig):
    def forward(self, preds, targets, protos, target_masks):
        """SegmentationLoss.forward (modules/segmentation_loss.py:26-75): the CUDA path covers overlap_masks=True with the
        BCE losses and no keypoints; everything else goes to the reference's own code."""
        model = self.model
        K = int(getattr(getattr(model, "proto_seg_module", None), "out_channels", 0) or 0)
        out_of_scope = (
            (self.alpha and self.gamma) or bool(getattr(model, "num_keypoints", None)) or not self.overlap_masks
            or self.batch_scale_loss or len(preds) != 3 or K not in (8, 16, 32)
            or not (isinstance(targets, torch.Tensor) and targets.dim() == 2 and targets.shape[1] == 6)
            or not all(_is_cuda_f32(p) for p in preds) or not _is_cuda_f32(protos)
            or not (isinstance(target_masks, torch.Tensor) and target_masks.dim() == 3 and target_masks.shape[0] == preds[0].shape[0])
            or type(self).loss_fn is not _saved.get("SegmentationLoss.loss_fn", type(self).loss_fn)
        )
        if out_of_scope:
            return orig(self, preds, targets, protos, target_masks)
        cfg = dict(anchor_t=self.anchor_t, edge_t=self.edge_t, box_w=self.box_w, conf_w=self.conf_w, class_w=self.class_w,
                   label_smoothing=self.label_smoothing, scale_w=self.scale_w, seg_w=self.seg_w)
        anchors3 = [model.sm_anchors.data, model.md_anchors.data, model.lg_anchors.data]
        dev = preds[0].device
        return ops.segmentation_loss([p if p.is_contiguous() else p.contiguous() for p in preds],
                                     targets.to(dev, torch.float32), protos if protos.is_contiguous() else protos.contiguous(),
                                     target_masks.to(dev), anchors3, cfg, model.num_classes, K)
    return forward


# ------------------------------------------------------------------------------------------------ B4
def _make_batched_nms(orig):
    def batched_nms(boxes, scores, idxs, iou_threshold):
        _need_cuda(boxes, "batched_nms")
        return ops.batched_nms(boxes, scores.float(), idxs.to(torch.int64), float(iou_threshold))
    return batched_nms


# ------------------------------------------------------------------------------------------------ B5
def _make_get_scale_pred(orig):
    def _get_scale_pred(self, scale_pred, anchors, input_shape, inference: bool = False):
        if self.num_keypoints is not None and self.num_keypoints > 0:
            return orig(self, scale_pred, anchors, input_shape, inference)
        if hasattr(self, "proto_seg_module"):
            # segmentation head: the detection decode plus tanh on the mask coefficients (detection.py:131-134)
            K = int(self.proto_seg_module.out_channels)
            if scale_pred.shape[-1] != self.num_classes + 5 + K or (inference and scale_pred.requires_grad and torch.is_grad_enabled()):
                return orig(self, scale_pred, anchors, input_shape, inference)
            _need_cuda(scale_pred, "_get_scale_pred")
            if not inference:     # training mode: one differentiable kernel each way (SegmentationLoss takes decoded rows)
                return ops.decode_train(scale_pred, self.num_classes, K)
            ishape = tuple(int(v) for v in input_shape)
            if _options["fuse_inference"] and "inference_seg.post_process_preds" in _saved:
                # stands for the decoded tensor of this scale (as for the detection model below); the wrapped
                # inference_seg.post_process_preds runs the fused decode+NMS on the head outputs
                raw = scale_pred if scale_pred.is_contiguous() else scale_pred.contiguous()
                return LazyPreds(tuple(raw.shape), [dict(raw=raw, anchors=anchors, input_shape=ishape, rescale=None, og_size=None,
                                                         num_classes=self.num_classes, tanh_cols=K)])
            return ops.decode_scale(scale_pred, anchors, ishape, True, None, self.num_classes, K)
        _need_cuda(scale_pred, "_get_scale_pred")
        if not inference:
            if scale_pred.shape[-1] != self.num_classes + 5:
                return orig(self, scale_pred, anchors, input_shape, inference)
            if _options["fuse_train_decode"]:
                # stands for the decoded tensor; DetectionLoss.forward takes the logits from it, any other consumer
                # gets the decoded values (differentiable CUDA decode) on first use
                if isinstance(scale_pred, LazyRows):
                    if scale_pred.pending and not scale_pred.decode:
                        return LazyRows(scale_pred.parts, decode=True)     # the head's three conv outputs, still apart
                    scale_pred = scale_pred.materialize()
                return LazyRows([scale_pred if scale_pred.is_contiguous() else scale_pred.contiguous()], decode=True)
            if isinstance(scale_pred, LazyRows):
                scale_pred = scale_pred.materialize()
            return ops.decode_train(scale_pred)
        if isinstance(scale_pred, LazyRows):
            scale_pred = scale_pred.materialize()
        ishape = tuple(int(v) for v in input_shape)
        if _options["fuse_inference"] and "post_process_preds" in _saved and scale_pred.shape[-1] == self.num_classes + 5:
            # stands for the decoded tensor of this scale; the wrapped post_process_preds runs the fused decode+NMS on
            # the head outputs, any other consumer gets the decoded values on first use
            raw = scale_pred if scale_pred.is_contiguous() else scale_pred.contiguous()
            return LazyPreds(tuple(raw.shape), [dict(raw=raw, anchors=anchors, input_shape=ishape, rescale=None, og_size=None,
                                                     num_classes=self.num_classes)])
        return ops.decode_scale(scale_pred, anchors, ishape, inference)
    return _get_scale_pred


_tls = threading.local()


def _make_net_forward(orig):
    def forward(self, x, inference: bool = False, og_size=None):
        # remembers og_size for _bbox_to_size, which only receives the derived device tensors (detection.py:77-81)
        _tls.og_size = tuple(int(v) for v in og_size) if og_size is not None else None
        try:
            return orig(self, x, inference, og_size)
        finally:
            _tls.og_size = None
    return forward


_detect_plans: Dict[Any, Any] = {}


def _make_post_process_preds(orig):
    def post_process_preds(imgs, preds, num_classes, colormap=None, iou_threshold: float = 0.5, score_threshold: float = 0.1,
                           vwriter=None, tracker=None, classmap=None, with_summary: bool = False, tracked_classes=None,
                           start_idx: int = 0, box_allowance=None):
        rest = dict(colormap=colormap, iou_threshold=iou_threshold, score_threshold=score_threshold, vwriter=vwriter,
                    tracker=tracker, classmap=classmap, with_summary=with_summary, tracked_classes=tracked_classes,
                    start_idx=start_idx, box_allowance=box_allowance)
        # 1. fused decode + score filter + per-image NMS on the head outputs (lines 57-89 of the reference function);
        # 2. the reference's own function on the kept candidates only
        small = _fused_rows(preds, int(num_classes), iou_threshold, score_threshold, box_allowance)
        if small is None:
            return orig(imgs, preds, num_classes, **rest)
        return orig(imgs, small, num_classes, **rest)
    return post_process_preds


def _fused_rows(preds, num_classes, iou_threshold, score_threshold, box_allowance):
    """Lines 57-89 of the reference's post_process_preds on the head outputs a pending LazyPreds stands for: fused decode +
    score filter + per-image NMS, then the kept candidates' rows of the decoded tensor, image by image, padded to a
    rectangle ``[B, kmax, D]`` with rows that can neither pass the score threshold nor suppress anything.  ``None`` if
    ``preds`` is not such a stand-in (the caller then runs the reference's function on what it was given)."""
    sc = preds.scales if isinstance(preds, LazyPreds) and preds.pending else None
    if sc is None or len(sc) != 3 or preds.dim() != 3 or not (score_threshold >= 0) \
            or len({(s["rescale"] is None, s["og_size"], s["input_shape"]) for s in sc}) != 1:
        return None
    raws = [s["raw"] for s in sc]
    og = sc[0]["og_size"] if sc[0]["rescale"] is not None else None
    key = (raws[0].device, tuple(tuple(r.shape) for r in raws), sc[0]["input_shape"], og, float(iou_threshold),
           float(score_threshold), box_allowance, int(num_classes))
    plan = _detect_plans.get(key)
    if plan is None:
        if len(_detect_plans) > 8:
            _detect_plans.clear()
        plan = _detect_plans[key] = ops.DetectPlan([tuple(r.shape) for r in raws], [s["anchors"] for s in sc], sc[0]["input_shape"],
                                                   int(num_classes), raws[0].device, og, float(iou_threshold),
                                                   float(score_threshold), box_allowance, None, "image")
    plan.enqueue(raws)
    det = plan.result()
    B, D = int(preds.shape[0]), int(preds.shape[2])
    counts = det.counts.to(torch.int64)
    kmax = max(int(counts.max()) if B else 0, 1)
    small = torch.zeros(B, kmax, D, dtype=torch.float32, device=raws[0].device)
    small[..., 0] = float("-inf")                                    # sigmoid(-inf) = 0: score 0, ranked last
    if det.keep_idxs.numel():
        rows = plan.decode_rows(det.keep_idxs)
        tc = int(sc[0].get("tanh_cols", 0))
        if tc:                                                       # the mask coefficients (detection.py:131-134)
            rows[:, 5 + num_classes: 5 + num_classes + tc].tanh_()
        offs = torch.zeros(B, dtype=torch.int64)
        offs[1:] = torch.cumsum(counts, 0)[:-1]
        offs = offs.to(rows.device, non_blocking=True)
        pos = torch.arange(rows.shape[0], device=rows.device) - offs[det.sample_idxs]
        small[det.sample_idxs, pos] = rows
    return small


def _make_seg_post_process_preds(orig):
    def post_process_preds(imgs, preds, protos, num_classes, colormap=None, iou_threshold: float = 0.5,
                           score_threshold: float = 0.1, vwriter=None, tracker=None, classmap=None, with_summary: bool = False,
                           tracked_classes=None, start_idx: int = 0, box_allowance=None):
        rest = dict(colormap=colormap, iou_threshold=iou_threshold, score_threshold=score_threshold, vwriter=vwriter,
                    tracker=tracker, classmap=classmap, with_summary=with_summary, tracked_classes=tracked_classes,
                    start_idx=start_idx, box_allowance=box_allowance)
        # 1. lines 57-89 fused on the head outputs when the model handed over a stand-in (install(DetectionNet=...,
        #    inference_seg=...)); otherwise they run as written (their torchvision.ops.batched_nms is the re-pointed one)
        small = _fused_rows(preds, int(num_classes), iou_threshold, score_threshold, box_allowance)
        if small is not None:
            preds = small
        # 2. the prototypes are marked so that the host loop's per-image ``sigmoid(coefs @ protos[i]) -> F.interpolate ->
        #    torch.gt(0.5)`` (lines 115-117) is recognised and runs on the two mask kernels (lazy.ProtoTrace / LazyMasks)
        if isinstance(protos, torch.Tensor) and protos.is_cuda and protos.dtype == torch.float32 and protos.dim() == 4 \
                and not protos.requires_grad and int(protos.shape[1]) <= 64:
            protos = protos.contiguous().as_subclass(ProtoTrace)
        return orig(imgs, preds, protos, num_classes, **rest)
    return post_process_preds


# ------------------------------------------------------------------------------------------------ f3
def _make_head_forward(orig):
    def forward(self, x):
        if hasattr(self, "masks_layer") or hasattr(self, "keypoints_layer") or not _options["split_head"] \
                or not (isinstance(x, torch.Tensor) and x.is_cuda):
            return orig(self, x)
        # the reference's own forward runs unchanged; only its final torch.cat([conf, cls, bbox], -1) (common.py:919)
        # is deferred: the three conv outputs stay where they are and the fused loss reads them in place
        out = orig(self, x.as_subclass(HeadTrace))
        return out if isinstance(out, LazyRows) else out.as_subclass(torch.Tensor)
    return forward


def _make_bbox_to_size(orig):
    def _bbox_to_size(self, pred, _from, _to):
        if self.num_keypoints is not None and self.num_keypoints > 0:
            return orig(self, pred, _from, _to)      # (keypoints are rescaled too, detection.py:186-189)
        _need_cuda(pred, "_bbox_to_size")
        if isinstance(pred, LazyPreds) and pred.pending:
            og = getattr(_tls, "og_size", None)
            if og is not None:
                out = pred.with_rescale(_from, _to)
                for sc in out.scales:
                    sc["og_size"] = og
                return out
        return ops.bbox_to_size(pred, _from, _to, self.num_classes)
    return _bbox_to_size


def _make_2dgrid(orig):
    cache: Dict[Any, torch.Tensor] = {}

    def _make_2dgrid(self, nx, ny, device="cpu"):
        # never reached under the patched _get_scale_pred (the kernels derive the cell from the row number); kept
        # callable for other users of the model, built once per (nx, ny, device) by the reference's own code
        key = (int(nx), int(ny), str(device))
        if key not in cache:
            cache[key] = orig(self, nx, ny, device)
        return cache[key].clone()
    return _make_2dgrid


# ------------------------------------------------------------------------------------------------ a13
def _make_ratio_metrics(orig, extras: bool):
    def ratio_metrics(anchors, wh_data, threshold: float = 4.0):
        if not (isinstance(wh_data, torch.Tensor) and wh_data.is_cuda):
            return orig(anchors, wh_data, threshold)   # host tensors (what train_det.py builds from the label files)
        wh = wh_data.to(torch.float32)
        if extras:
            return ops.ratio_metrics_w_extras(anchors, wh, threshold)
        return ops.ratio_metrics(anchors, wh, threshold)
    ratio_metrics.__name__ = orig.__name__
    return ratio_metrics


_options = {"fuse_train_decode": True, "split_head": True, "fuse_inference": True}


def install(DetectionDataset=None, DetectionLoss=None, DetectionNet=None, torchvision_ops=True, make_anchors=None,
            fuse_train_decode: bool = True, EffiDecHead=None, split_head: bool = True, inference_det=None,
            fuse_inference: bool = True, SegmentationLoss=None, inference_seg=None) -> None:
    """Re-point the reference's call sites at the CUDA operators.  Pass the reference classes that are
    imported in your process (any subset); ``torchvision_ops=True`` also replaces
    ``torchvision.ops.batched_nms`` (what ``inference_det.py:77`` looks up at call time); ``make_anchors`` is the
    reference's ``utils.make_anchors`` module (``ratio_metrics*``).  ``fuse_train_decode=False`` makes the
    training-mode ``_get_scale_pred`` return a real decoded tensor (one CUDA kernel each way) instead of the
    deferred stand-in.  ``EffiDecHead`` (modules/common.py:852-931): its final ``torch.cat`` is deferred as well, so the
    loss reads the head's three conv outputs in place (zero-copy when the model runs channels-last,
    otherwise the pieces are made contiguous -- the copy the concatenation would have been).  ``SegmentationLoss``
    (modules/segmentation_loss.py): its ``forward`` runs the fused detection terms plus the mask-term kernels
    (overlap_masks=True, BCE, no keypoints -- other configurations keep the reference's code).
    ``inference_seg`` (the reference's script module): its ``post_process_preds`` builds the boolean masks of the kept rows
    with the mask kernels (inference_seg.py:115-117)."""
    _options["fuse_train_decode"] = bool(fuse_train_decode)
    _options["split_head"] = bool(split_head)
    _options["fuse_inference"] = bool(fuse_inference)
    if inference_det is not None and "post_process_preds" not in _saved:
        # ``inference_det`` (the reference's script module): evaluate_frames looks post_process_preds up in the module's
        # globals at call time (inference_det.py:199-207,229-239)
        _saved["post_process_preds"] = (inference_det, inference_det.post_process_preds)
        inference_det.post_process_preds = _make_post_process_preds(inference_det.post_process_preds)
    if inference_seg is not None and "inference_seg.post_process_preds" not in _saved:
        _saved["inference_seg.post_process_preds"] = (inference_seg, inference_seg.post_process_preds)
        inference_seg.post_process_preds = _make_seg_post_process_preds(inference_seg.post_process_preds)
    if EffiDecHead is not None and "EffiDecHead.forward" not in _saved:
        _saved["EffiDecHead.forward"] = (EffiDecHead, EffiDecHead.__dict__["forward"])
        EffiDecHead.forward = _make_head_forward(EffiDecHead.__dict__["forward"])
    from . import _lib
    _lib.lib()  # fail loudly now if the extension is not built
    if DetectionDataset is not None and "build_target_by_scale" not in _saved:
        _saved["build_target_by_scale"] = (DetectionDataset, DetectionDataset.__dict__["build_target_by_scale"])
        DetectionDataset.build_target_by_scale = staticmethod(
            _make_build_target_by_scale(DetectionDataset.build_target_by_scale))
    if DetectionLoss is not None and "compute_ciou" not in _saved:
        _saved["compute_ciou"] = (DetectionLoss, DetectionLoss.__dict__["compute_ciou"])
        _saved["forward"] = (DetectionLoss, DetectionLoss.__dict__["forward"])
        _saved["DetectionLoss.loss_fn"] = DetectionLoss.__dict__["loss_fn"]
        DetectionLoss.compute_ciou = staticmethod(_make_compute_ciou(DetectionLoss.compute_ciou))
        DetectionLoss.forward = _make_loss_forward(DetectionLoss.__dict__["forward"])
    if DetectionNet is not None and "_get_scale_pred" not in _saved:
        _saved["_get_scale_pred"] = (DetectionNet, DetectionNet.__dict__["_get_scale_pred"])
        DetectionNet._get_scale_pred = _make_get_scale_pred(DetectionNet.__dict__["_get_scale_pred"])
        if "forward" in DetectionNet.__dict__:
            _saved["DetectionNet.forward"] = (DetectionNet, DetectionNet.__dict__["forward"])
            DetectionNet.forward = _make_net_forward(DetectionNet.__dict__["forward"])
        if "_bbox_to_size" in DetectionNet.__dict__:
            _saved["_bbox_to_size"] = (DetectionNet, DetectionNet.__dict__["_bbox_to_size"])
            DetectionNet._bbox_to_size = _make_bbox_to_size(DetectionNet.__dict__["_bbox_to_size"])
        if "_make_2dgrid" in DetectionNet.__dict__:
            _saved["_make_2dgrid"] = (DetectionNet, DetectionNet.__dict__["_make_2dgrid"])
            DetectionNet._make_2dgrid = _make_2dgrid(DetectionNet.__dict__["_make_2dgrid"])
    if SegmentationLoss is not None and "SegmentationLoss.forward" not in _saved:
        _saved["SegmentationLoss.forward"] = (SegmentationLoss, SegmentationLoss.__dict__["forward"])
        _saved["SegmentationLoss.loss_fn"] = SegmentationLoss.__dict__["loss_fn"]
        SegmentationLoss.forward = _make_seg_loss_forward(SegmentationLoss.__dict__["forward"])
    if make_anchors is not None and "ratio_metrics" not in _saved:
        _saved["ratio_metrics"] = (make_anchors, make_anchors.ratio_metrics)
        _saved["ratio_metrics_w_extras"] = (make_anchors, make_anchors.ratio_metrics_w_extras)
        make_anchors.ratio_metrics = _make_ratio_metrics(make_anchors.ratio_metrics, False)
        make_anchors.ratio_metrics_w_extras = _make_ratio_metrics(make_anchors.ratio_metrics_w_extras, True)
    if torchvision_ops and "batched_nms" not in _saved:
        import torchvision
        _saved["batched_nms"] = (torchvision.ops, torchvision.ops.batched_nms)
        torchvision.ops.batched_nms = _make_batched_nms(torchvision.ops.batched_nms)


_NAMES = ("build_target_by_scale", "compute_ciou", "forward", "_get_scale_pred", "_bbox_to_size", "_make_2dgrid",
          "batched_nms", "ratio_metrics", "ratio_metrics_w_extras", "post_process_preds")


def uninstall() -> None:
    for name in _NAMES:
        if name in _saved:
            owner, orig = _saved.pop(name)
            setattr(owner, name, orig)
    for key in ("EffiDecHead.forward", "DetectionNet.forward", "SegmentationLoss.forward"):
        if key in _saved:
            owner, orig = _saved.pop(key)
            owner.forward = orig
    if "inference_seg.post_process_preds" in _saved:
        owner, orig = _saved.pop("inference_seg.post_process_preds")
        owner.post_process_preds = orig
    _detect_plans.clear()
    _saved.pop("DetectionLoss.loss_fn", None)
    _saved.pop("SegmentationLoss.loss_fn", None)


def installed() -> Dict[str, bool]:
    return {k: (k in _saved) for k in _NAMES}
