"""ctypes front-end of the CPU oracle (``oracle/boxgeom_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``boxgeom_oracle.c``.  Nothing under
``vision_conglomerate_b200/`` imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs do, and only as the checker / the timed CPU baseline.

The functions mirror the reference routines they restate (cited per function in
the C file) and work on numpy arrays (torch CPU tensors are accepted and
converted).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libboxgeom_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "boxgeom_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libboxgeom_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, i64p, i32p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_double))
        L.bgo_decode_scale.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int,
                                       C.c_int, f32p]
        L.bgo_decode_scale.restype = None
        L.bgo_bbox_to_size.argtypes = [f32p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.bgo_bbox_to_size.restype = None
        L.bgo_score_xyxy.argtypes = [f32p, C.c_size_t, C.c_int, C.c_float, f32p, i32p, f32p]
        L.bgo_score_xyxy.restype = None
        L.bgo_nms.argtypes = [f32p, f32p, C.c_int64, C.c_double, i64p]
        L.bgo_nms.restype = C.c_int64
        L.bgo_batched_nms.argtypes = [f32p, f32p, i64p, C.c_int64, C.c_double, i64p]
        L.bgo_batched_nms.restype = C.c_int64
        L.bgo_assign.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, f32p, C.c_int, C.c_float, C.c_float, C.c_int64,
                                 i64p, i64p, f32p, f32p]
        L.bgo_assign.restype = C.c_int64
        L.bgo_assign_ex.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_float, C.c_float, C.c_int64,
                                    i64p, i64p, f32p, f32p, i64p]
        L.bgo_assign_ex.restype = C.c_int64
        L.bgo_ciou.argtypes = [f32p, f32p, C.c_int64, C.c_float, f32p, f32p]
        L.bgo_ciou.restype = None
        L.bgo_loss_scale.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64p, C.c_int64, i64p, f32p,
                                     f32p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, f64p, i64p, f32p]
        L.bgo_loss_scale.restype = None
        L.bgo_ratio_metrics.argtypes = [f32p, C.c_int64, f32p, C.c_int, C.c_float, f64p]
        L.bgo_ratio_metrics.restype = None
        _lib = L
    return _lib


def _np(x, dtype) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


# ------------------------------------------------------------------ inference side
def decode_scale(raw, anchors, H: int, W: int, inference: bool = True) -> np.ndarray:
    raw = _np(raw, np.float32)
    anchors = _np(anchors, np.float32)
    B, ny, nx, na, D = raw.shape
    out = np.empty_like(raw)
    lib().bgo_decode_scale(_p(raw, C.c_float), B, ny, nx, na, D - 5, _p(anchors, C.c_float), H, W,
                           int(inference), _p(out, C.c_float))
    return out


def decode_inference(raws: Sequence, anchors3: Sequence, H: int, W: int,
                     og_size: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """modules/detection.py:58-91 with ``inference=True`` from the three head outputs on:
    decode each scale, optional rescale (guard at :76), reshape and concatenate to ``[B, N, 5+C]``."""
    outs = []
    for raw, anc in zip(raws, anchors3):
        d = decode_scale(raw, anc, H, W, True)
        B, D = d.shape[0], d.shape[-1]
        d = d.reshape(B, -1, D)
        if og_size is not None and (og_size[0] != H and og_size[1] != W):
            flat = d.reshape(-1, D)
            lib().bgo_bbox_to_size(_p(flat, C.c_float), flat.shape[0], D - 5, H, W, int(og_size[0]), int(og_size[1]))
        outs.append(d)
    return np.concatenate(outs, axis=1)


def score_xyxy(preds, box_allowance: float = 0.0):
    preds = _np(preds, np.float32)
    D = preds.shape[-1]
    flat = preds.reshape(-1, D)
    n = flat.shape[0]
    score = np.empty(n, np.float32)
    cls = np.empty(n, np.int32)
    xyxy = np.empty((n, 4), np.float32)
    lib().bgo_score_xyxy(_p(flat, C.c_float), n, D - 5, float(box_allowance or 0.0), _p(score, C.c_float),
                         _p(cls, C.c_int32), _p(xyxy, C.c_float))
    return score, cls, xyxy


def nms(boxes, scores, iou_threshold: float) -> np.ndarray:
    boxes, scores = _np(boxes, np.float32), _np(scores, np.float32)
    n = scores.shape[0]
    keep = np.empty(max(n, 1), np.int64)
    k = lib().bgo_nms(_p(boxes, C.c_float), _p(scores, C.c_float), n, float(iou_threshold), _p(keep, C.c_int64))
    return keep[:k].copy()


def batched_nms(boxes, scores, idxs, iou_threshold: float) -> np.ndarray:
    """Canonical order: score descending, index ascending inside equal scores."""
    boxes, scores, idxs = _np(boxes, np.float32), _np(scores, np.float32), _np(idxs, np.int64)
    n = scores.shape[0]
    keep = np.empty(max(n, 1), np.int64)
    k = lib().bgo_batched_nms(_p(boxes, C.c_float), _p(scores, C.c_float), _p(idxs, C.c_int64), n,
                              float(iou_threshold), _p(keep, C.c_int64))
    return keep[:k].copy()


def canonical_keep(keep, scores) -> np.ndarray:
    """Canonicalise a torchvision-ordered keep list to (score desc, index asc)."""
    keep = _np(keep, np.int64)
    s = _np(scores, np.float32)[keep]
    order = np.lexsort((keep, -s.astype(np.float64)))
    return keep[order]


def post_process(preds, iou_threshold: float, score_threshold: float, box_allowance: Optional[float] = None,
                 tracked_classes: Optional[Sequence[int]] = None) -> Dict[str, np.ndarray]:
    """inference_det.py:57-97 (+107-109): returns ``pred_boxes [K',6]`` = (score, cls, x1,y1,x2,y2),
    ``sample_idxs [K']`` and ``keep`` (flat indices, canonical order) after the strict score threshold
    and the optional tracked-class row filter."""
    preds = _np(preds, np.float32)
    B, N, D = preds.shape
    score, cls, xyxy = score_xyxy(preds, box_allowance or 0.0)
    sample = np.repeat(np.arange(B, dtype=np.int64), N)
    keep = batched_nms(xyxy, score, sample, iou_threshold)
    keep = keep[score[keep] > np.float32(score_threshold)]
    if tracked_classes:
        keep = keep[np.isin(cls[keep], np.asarray(tracked_classes))]
    pb = np.concatenate([score[keep, None], cls[keep, None].astype(np.float32), xyxy[keep]], axis=1)
    return {"pred_boxes": pb, "sample_idxs": sample[keep], "keep": keep, "score": score, "cls": cls, "xyxy": xyxy}


# ------------------------------------------------------------------- training side
def build_target_by_scale(targets, fmap_shape, anchors, anchor_threshold: float = 4.0,
                          edge_threshold: float = 0.5):
    targets, anchors = _np(targets, np.float32).reshape(-1, 6), _np(anchors, np.float32)
    nt, na = targets.shape[0], anchors.shape[0]
    ny, nx = int(fmap_shape[0]), int(fmap_shape[1])
    cap = max(5 * na * nt, 1)
    idx4 = np.empty((4, cap), np.int64)
    cls = np.empty(cap, np.int64)
    anc = np.empty((cap, 2), np.float32)
    box = np.empty((cap, 4), np.float32)
    M = lib().bgo_assign(_p(targets, C.c_float), nt, ny, nx, _p(anchors, C.c_float), na, float(anchor_threshold),
                         float(edge_threshold), cap, _p(idx4, C.c_int64), _p(cls, C.c_int64), _p(anc, C.c_float),
                         _p(box, C.c_float))
    assert M >= 0
    return [idx4[k, :M].copy() for k in range(4)], cls[:M].copy(), anc[:M].copy(), box[:M].copy()


def build_target_by_scale_ex(targets, fmap_shape, anchors, anchor_threshold: float = 4.0, edge_threshold: float = 0.5,
                             overlap_masks: Optional[bool] = None, batch_size: Optional[int] = None):
    """dataset/detection_dataset.py:90-246 including the segmentation (``overlap_masks``) and keypoint-column
    variants.  Returns (indices, classes, anchors, boxes, tmask_idx or None, keypoints or None)."""
    targets, anchors = _np(targets, np.float32), _np(anchors, np.float32)
    targets = targets.reshape(-1, targets.shape[-1] if targets.ndim == 2 else 6)
    nt, stride, na = targets.shape[0], targets.shape[1], anchors.shape[0]
    ny, nx = int(fmap_shape[0]), int(fmap_shape[1])
    cap = max(5 * na * nt, 1)
    idx4 = np.empty((4, cap), np.int64)
    cls = np.empty(cap, np.int64)
    anc = np.empty((cap, 2), np.float32)
    box = np.empty((cap, 4), np.float32)
    src = np.empty(cap, np.int64)
    M = lib().bgo_assign_ex(_p(targets, C.c_float), nt, stride, ny, nx, _p(anchors, C.c_float), na, float(anchor_threshold),
                            float(edge_threshold), cap, _p(idx4, C.c_int64), _p(cls, C.c_int64), _p(anc, C.c_float),
                            _p(box, C.c_float), _p(src, C.c_int64))
    assert M >= 0
    src = src[:M]
    tmask = None
    if overlap_masks is not None:
        if overlap_masks:  # :146-157: 1 + position inside the image's block, blocks sized by the per-image counts
            if not batch_size:
                raise ValueError("batch_size is required when overlap_mask is set to True")
            per_t = np.concatenate([np.arange(int((targets[:, 0] == i).sum())) + 1 for i in range(batch_size)]) \
                if nt else np.zeros(0, np.int64)
            if per_t.shape[0] != nt:
                raise ValueError("per-image target counts do not add up to the number of targets")
        else:              # :170: the target's own position
            per_t = np.arange(nt)
        tmask = per_t[src].astype(np.int64)
    kpts = targets[src, 6:].copy() if stride > 6 else None
    return [idx4[k, :M].copy() for k in range(4)], cls[:M].copy(), anc[:M].copy(), box[:M].copy(), tmask, kpts


def compute_ciou(p, t, e: float = 1e-7, with_grad: bool = False):
    p, t = _np(p, np.float32).reshape(-1, 4), _np(t, np.float32).reshape(-1, 4)
    M = p.shape[0]
    out = np.empty(M, np.float32)
    g = np.empty((M, 4), np.float32) if with_grad else None
    lib().bgo_ciou(_p(p, C.c_float), _p(t, C.c_float), M, float(e), _p(out, C.c_float),
                   _p(g, C.c_float) if with_grad else None)
    return (out, g) if with_grad else out


def macro_metrics(hist: np.ndarray, M: int) -> Dict[str, float]:
    """sklearn accuracy / macro f1 / precision / recall from per-class (tp, n_true, n_pred) counts
    (modules/detection_loss.py:198-206)."""
    if M == 0:
        return dict(accuracy=float("nan"), f1=float("nan"), precision=float("nan"), recall=float("nan"))
    tp, nt, npred = (hist[i].astype(np.float64) for i in range(3))
    lab = (nt + npred) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = np.where(npred > 0, tp / npred, 0.0)[lab]
        rec = np.where(nt > 0, tp / nt, 0.0)[lab]
        f1 = (2 * tp / (nt + npred))[lab]
    return dict(accuracy=float(tp.sum() / M), f1=float(f1.mean()), precision=float(prec.mean()),
                recall=float(rec.mean()))


def loss_scale(preds, targets, anchors, cfg: dict, w_scale: float, with_grad: bool = False):
    preds = _np(preds, np.float32)
    B, ny, nx, na, D = preds.shape
    Cc = D - 5
    idx, cls, anc, box = build_target_by_scale(targets, (ny, nx), anchors, cfg["anchor_t"], cfg["edge_t"])
    M = cls.shape[0]
    cap = max(M, 1)
    idx4 = np.zeros((4, cap), np.int64)
    for k in range(4):
        idx4[k, :M] = idx[k]
    scal = np.zeros(8, np.float64)
    hist = np.zeros((3, Cc), np.int64)
    grad = np.empty_like(preds) if with_grad else None
    lib().bgo_loss_scale(_p(preds, C.c_float), B, ny, nx, na, Cc, _p(idx4, C.c_int64), cap,
                         _p(np.ascontiguousarray(cls), C.c_int64), _p(np.ascontiguousarray(anc), C.c_float),
                         _p(np.ascontiguousarray(box), C.c_float), M, float(cfg["label_smoothing"]),
                         float(cfg["box_w"] * w_scale), float(cfg["conf_w"] * w_scale),
                         float(cfg["class_w"] * w_scale), _p(scal, C.c_double), _p(hist, C.c_int64),
                         _p(grad, C.c_float) if with_grad else None)
    m = dict(mean_ciou=scal[3], conf_loss=scal[1], avg_pos_conf=scal[4], avg_neg_conf=scal[5],
             class_loss=scal[2] if M else float("nan"))
    m.update(macro_metrics(hist, M))
    return scal, m, grad, M


def train_decode(raw) -> np.ndarray:
    """Training-mode ``DetectionNet._get_scale_pred`` (modules/detection.py:122,125,164): xy = 2s - 0.5,
    wh = (2s)^2 on the four box columns, objectness / class logits copied."""
    raw = _np(raw, np.float32)
    return decode_scale(raw, np.ones((raw.shape[3], 2), np.float32), 1, 1, inference=False)


def train_decode_backward(raw, grad_decoded) -> np.ndarray:
    """Chain rule of :func:`train_decode` (what autograd does through sigmoid / mul / sub / pow), in float64."""
    raw = _np(raw, np.float32)
    g = np.array(grad_decoded, dtype=np.float64, copy=True)
    Cc = raw.shape[-1] - 5
    s = 1.0 / (1.0 + np.exp(-raw[..., Cc + 1:Cc + 5].astype(np.float64)))
    g[..., Cc + 1:Cc + 3] *= 2.0 * s[..., :2] * (1.0 - s[..., :2])
    g[..., Cc + 3:Cc + 5] *= 8.0 * s[..., 2:] ** 2 * (1.0 - s[..., 2:])
    return g.astype(np.float32)


def detection_loss(preds3: Sequence, targets, anchors3: Sequence, cfg: dict, with_grad: bool = False,
                   input_form: str = "decoded"):
    """modules/detection_loss.py:84-122 (DetectionLoss.forward) on the default BCE configuration.
    Returns (loss, metrics_dict, [grad_sm, grad_md, grad_lg] or None, [M_sm, M_md, M_lg]).

    ``input_form="raw"``: ``preds3`` are the head's own outputs; the training-mode decode of
    modules/detection.py:98-173 runs first (what ``DetectionNet.forward`` does before the loss sees the tensors)
    and the gradients are taken back through it.  ``"split"``: per scale ``(conf [B,ny,nx,na], cls [...,C],
    bbox [...,4])``, the conv outputs ``EffiDecHead.forward`` concatenates (modules/common.py:908-919); gradients
    come back as the same triples."""
    raws = None
    if input_form == "split":
        raws = [np.concatenate([_np(c, np.float32).reshape(*_np(k, np.float32).shape[:4], 1), _np(k, np.float32),
                                _np(b, np.float32)], axis=-1) for c, k, b in preds3]
    elif input_form == "raw":
        raws = [_np(p, np.float32) for p in preds3]
    if raws is not None:
        preds3 = [train_decode(r) for r in raws]
    sw = cfg.get("scale_w") or [4.0, 2.0, 1.0]
    lbox = lconf = lcls = 0.0
    rows, grads, Ms = [], [], []
    for p, a, w in zip(preds3, anchors3, sw):
        scal, m, g, M = loss_scale(p, targets, a, cfg, w, with_grad)
        lbox += w * scal[0]
        lconf += w * scal[1]
        lcls += w * scal[2]
        rows.append(m)
        grads.append(g)
        Ms.append(M)
    loss = cfg["box_w"] * lbox + cfg["conf_w"] * lconf + cfg["class_w"] * lcls
    metrics = {"aggregate_loss": float(loss)}
    for k in rows[0]:
        vals = np.array([r[k] for r in rows], np.float64)
        metrics[k] = float(np.nanmean(vals)) if not np.all(np.isnan(vals)) else float("nan")
    if with_grad and raws is not None:
        grads = [train_decode_backward(r, g) for r, g in zip(raws, grads)]
        if input_form == "split":
            grads = [(g[..., 0].copy(), g[..., 1:-4].copy(), g[..., -4:].copy()) for g in grads]
    return float(loss), metrics, (grads if with_grad else None), Ms


def ratio_metrics(anchors, wh, threshold: float = 4.0) -> Tuple[float, float, float]:
    anchors, wh = _np(anchors, np.float32).reshape(-1, 2), _np(wh, np.float32).reshape(-1, 2)
    out = np.zeros(3, np.float64)
    lib().bgo_ratio_metrics(_p(wh, C.c_float), wh.shape[0], _p(anchors, C.c_float), anchors.shape[0],
                            float(threshold), _p(out, C.c_double))
    return float(out[0]), float(out[1]), float(out[2])
