"""Import the *unmodified* reference: from /root/reference in the build container, or from the untouched copy
``baseline/_ref/`` that ``__graft_entry__.build()`` makes there (git-ignored, but it travels to the GPU box with
the snapshot; the reference is a script tree without setup.py / pyproject.toml, so "installing" it is a copy).

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/``,
by the tests that run the drop-in layer on the REAL reference classes and compare patched against unpatched
results, and by ``bench.py``'s reference arm / ``gpu_library_baseline`` leg (the reference's own code on the
host cores and on torch-CUDA / torchvision-CUDA).  Never imported by the product.

Recipe from SURVEY.md Appendix B: the reference imports ``supervision`` at module
top (utils/utils.py:11, inference_det.py:12), which is not installed, so a stub
module is registered first.  ``DetectionLoss`` only needs ``model.{sm,md,lg}_anchors``
and ``model.num_classes`` (modules/detection_loss.py:91-93,141) -> ``FakeModel``.
"""
from __future__ import annotations

import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find():
    for cand in (os.environ.get("BOXGEOM_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "modules", "detection_loss.py")):
            return cand
    return os.environ.get("BOXGEOM_REFERENCE", "/root/reference")


REF = _find()


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "modules", "detection_loss.py"))


def install_copy(dst=None) -> str:
    """Copy the reference's source tree (code and configs; not its 4 MB of PNG/PDF resources) to baseline/_ref."""
    import shutil
    dst = dst or os.path.join(_ROOT, "baseline", "_ref")
    src = "/root/reference"
    if not os.path.isfile(os.path.join(src, "modules", "detection_loss.py")):
        raise RuntimeError("no reference at %s" % src)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("resources", ".git", "__pycache__", "*.pyc"))
    return dst


_mods = None


def load():
    """Returns a namespace with DetectionNet, DetectionLoss, DetectionDataset, inference_det,
    make_anchors, utils and FakeModel."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference not present at %s" % REF)
    import torch
    import torch.nn as nn

    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "supervision" not in sys.modules:
        sv = types.ModuleType("supervision")
        sv.Detections = type("Detections", (), {})
        sv.ByteTrack = type("ByteTrack", (), {"__init__": lambda s, *a, **k: None})
        sys.modules["supervision"] = sv
    from modules.detection import DetectionNet
    from modules.common import EffiDecHead
    from modules.detection_loss import DetectionLoss
    from dataset.detection_dataset import DetectionDataset
    import inference_det
    from utils import make_anchors, utils as ref_utils

    inference_det.device = "cpu"  # module-global only defined under __main__ (inference_det.py:108 vs :318)

    class FakeModel(nn.Module):
        def __init__(self, num_classes, anchors):
            super().__init__()
            self.num_classes, self.num_keypoints = num_classes, 0
            for k in ("sm", "md", "lg"):
                setattr(self, k + "_anchors", nn.Parameter(torch.tensor(anchors[k], dtype=torch.float32)))

    class DecodeOnly:
        """Carries just what DetectionNet._get_scale_pred / _bbox_to_size / _make_2dgrid read from self."""
        num_keypoints = None

        def __init__(self, num_classes):
            self.num_classes = num_classes

        _get_scale_pred = DetectionNet._get_scale_pred
        _bbox_to_size = DetectionNet._bbox_to_size
        _make_2dgrid = DetectionNet._make_2dgrid

    from modules.segmentation_loss import SegmentationLoss
    from modules.segmentation import SegmentationNet

    class FakeSegModel(FakeModel):
        """What SegmentationLoss reads from its model on top of FakeModel: proto_seg_module.out_channels
        (modules/segmentation_loss.py:101)."""
        def __init__(self, num_classes, anchors, num_masks):
            super().__init__(num_classes, anchors)
            self.proto_seg_module = types.SimpleNamespace(out_channels=num_masks)

    ns = types.SimpleNamespace(SegmentationLoss=SegmentationLoss, FakeSegModel=FakeSegModel, SegmentationNet=SegmentationNet,
                               DetectionNet=DetectionNet, DetectionLoss=DetectionLoss, EffiDecHead=EffiDecHead,
                               DetectionDataset=DetectionDataset, inference_det=inference_det,
                               make_anchors=make_anchors, utils=ref_utils, FakeModel=FakeModel,
                               DecodeOnly=DecodeOnly)
    _mods = ns
    return ns


def model_config(task: str = "detection") -> dict:
    """``model_config`` of the reference's config/{detection,segmentation}/config.yaml (CSPBackBone + RepBiPAN + EffiDecHead)."""
    import yaml
    with open(os.path.join(REF, "config", task, "config.yaml")) as f:
        return yaml.safe_load(f)["model_config"]


def ref_decode_inference(raws, anchors3, H, W, og_size=None, num_classes=80):
    """Runs the reference's own _get_scale_pred x3, the rescale guard and the reshape/cat of
    DetectionNet.forward (modules/detection.py:69-91) starting from the three head outputs."""
    import torch
    ns = load()
    m = ns.DecodeOnly(num_classes)
    with torch.no_grad():
        ps = [m._get_scale_pred(r.clone(), a, input_shape=(H, W), inference=True) for r, a in zip(raws, anchors3)]
        if (og_size is not None) and (og_size[0] != H and og_size[1] != W):
            _from = torch.tensor([W, H, W, H])
            _to = torch.tensor([og_size[1], og_size[0], og_size[1], og_size[0]])
            ps = [m._bbox_to_size(p, _from, _to) for p in ps]
        B = raws[0].shape[0]
        D = raws[0].shape[-1]
        ps = [p.reshape(B, -1, D) for p in ps]
        return torch.cat(ps, dim=1).flatten(start_dim=1, end_dim=-2)


def ref_post_process(preds, num_classes, iou_threshold, score_threshold, box_allowance=None, tracked_classes=None):
    """Calls the reference's post_process_preds unmodified and captures (a) the arguments and result of
    its torchvision.ops.batched_nms call and (b) the per-image box arrays it hands to the drawing code
    (after the score threshold and the tracked-class filter)."""
    import numpy as np
    import torch
    import torchvision
    ns = load()
    inf = ns.inference_det
    cap = {"per_image": []}
    real_nms = torchvision.ops.batched_nms

    def spy_nms(boxes, scores, idxs, iou_threshold):
        keep = real_nms(boxes, scores, idxs, iou_threshold)
        cap.update(boxes=boxes.clone(), scores=scores.clone(), idxs=idxs.clone(), keep=keep.clone())
        return keep

    def spy_apply(img, boxes, **kw):
        cap["per_image"].append(np.array(boxes, copy=True))
        return img

    class _NullImg:
        @staticmethod
        def fromarray(a):
            class _I:
                def save(self, f):
                    pass
            return _I()

    old = (torchvision.ops.batched_nms, inf.apply_bboxes, inf.Image, inf.STORAGE_PATH)
    import tempfile
    tmp = tempfile.mkdtemp()
    try:
        torchvision.ops.batched_nms = spy_nms
        inf.apply_bboxes = spy_apply
        inf.Image = _NullImg
        inf.STORAGE_PATH = tmp
        B = preds.shape[0]
        imgs = torch.zeros(B, 3, 8, 8, dtype=torch.uint8, device=preds.device)
        with torch.no_grad():
            # (the reference adds the box allowance in place on a view of preds: hand it a copy of a real tensor; a
            # stand-in of the drop-in layer is passed as evaluate_frames passes it -- untouched)
            inf.post_process_preds(imgs, preds.clone() if type(preds) is torch.Tensor else preds, num_classes, iou_threshold=iou_threshold,
                                   score_threshold=score_threshold, box_allowance=box_allowance,
                                   tracked_classes=list(tracked_classes) if tracked_classes else None)
    finally:
        torchvision.ops.batched_nms, inf.apply_bboxes, inf.Image, inf.STORAGE_PATH = old
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return cap


def ref_seg_post_process(preds, protos, num_classes, iou_threshold, score_threshold, box_allowance=None,
                         tracked_classes=None, img_size=None, capture_values=False):
    """Calls the reference's inference_seg.post_process_preds unmodified (inference_seg.py:40-175) and captures the
    arguments / result of its torchvision.ops.batched_nms call and, per surviving image, the box array, the boolean
    masks and the keypoint array it hands to the drawing code."""
    import numpy as np
    import torch
    import torchvision
    load()
    import inference_seg as inf
    inf.device = preds.device  # module-global only defined under __main__, read by the tracked-class filter
    cap = {"per_image": [], "masks": [], "keypoints": []}
    real_nms = torchvision.ops.batched_nms

    def spy_nms(boxes, scores, idxs, iou_threshold):
        keep = real_nms(boxes, scores, idxs, iou_threshold)
        cap.update(boxes=boxes.clone(), scores=scores.clone(), idxs=idxs.clone(), keep=keep.clone())
        return keep

    def spy_boxes(img, boxes, **kw):
        cap["per_image"].append(np.array(boxes, copy=True))
        return img

    def spy_segments(img, masks, **kw):
        cap["masks"].append(np.array(masks, copy=True))
        return img

    def spy_keypoints(img, kp, **kw):
        cap["keypoints"].append(np.array(kp, copy=True))
        return img

    class _NullImg:
        @staticmethod
        def fromarray(a):
            class _I:
                def save(self, f):
                    pass
            return _I()

    old_F = inf.F
    if capture_values:  # the interpolated values behind the boolean masks (inference_seg.py:116), for threshold-band checks
        cap["values"] = []

        def spy_interpolate(x, *a, **k):
            out = old_F.interpolate(x, *a, **k)
            cap["values"].append(out[0].detach().cpu().numpy())
            return out
        inf.F = types.SimpleNamespace(interpolate=spy_interpolate)
    old = (torchvision.ops.batched_nms, inf.apply_bboxes, inf.apply_segments, inf.apply_keypoints, inf.Image, inf.STORAGE_PATH)
    import tempfile
    tmp = tempfile.mkdtemp()
    try:
        torchvision.ops.batched_nms = spy_nms
        inf.apply_bboxes, inf.apply_segments, inf.apply_keypoints = spy_boxes, spy_segments, spy_keypoints
        inf.Image = _NullImg
        inf.STORAGE_PATH = tmp
        B = preds.shape[0]
        # same size as the protos (the bilinear resize of inference_seg.py:116 is then the identity) unless img_size is given
        ih, iw = img_size if img_size is not None else (protos.shape[2], protos.shape[3])
        imgs = torch.zeros(B, 3, ih, iw, dtype=torch.uint8)
        with torch.no_grad():
            # (the reference adds the box allowance in place; a stand-in of the drop-in is handed over as it is)
            inf.post_process_preds(imgs, preds.clone() if type(preds) is torch.Tensor else preds, protos.clone(), num_classes,
                                   iou_threshold=iou_threshold,
                                   score_threshold=score_threshold, box_allowance=box_allowance,
                                   tracked_classes=list(tracked_classes) if tracked_classes else None)
    finally:
        (torchvision.ops.batched_nms, inf.apply_bboxes, inf.apply_segments, inf.apply_keypoints, inf.Image, inf.STORAGE_PATH) = old
        inf.F = old_F
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return cap
