"""Keep-list parity accounting between the CUDA path and the CPU oracle.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg as the checker).

``north_star`` asks for bit-exact NMS keep-lists.  With *identical fp32 boxes and scores* the CUDA NMS is
bit-exact (tests: ``test_nms_*``).  In the fused decode+NMS path the boxes and scores are themselves computed on
the device: ``expf`` of the CUDA math library, glibc's and ATen's differ by an ulp on some arguments, so a score
or a box coordinate can differ in its last bit, and a greedy decision that sits *exactly* on a threshold can flip.
:func:`explain_keep_mismatches` does not tolerate a count; it demands that every candidate on which the two lists
disagree is one of these provably marginal cases and returns what it found, so that tests and the bench line
print the measured numbers:

* ``score_threshold``  the candidate's score is within ``score_ulps`` ulps of the score threshold;
* ``iou_at_threshold`` a deciding pair -- the candidate and a box kept by either list (or itself in dispute)
                       -- has an IoU within ``iou_band`` of the IoU threshold.  ``iou_band`` = 2e-6 by default:
                       one ulp of a sigmoid moves a coordinate by at most stride*2*6e-8 and an extent by
                       1.2e-7 relative, i.e. the IoU by ~5e-7 (four extents enter it);
* ``score_tie``        an overlapping box (IoU above the threshold) has a score within ``score_ulps`` ulps of the
                       candidate's: the two implementations may rank the pair differently;
* ``cascade``          the candidate overlaps (IoU above the threshold) a higher-ranked candidate already
                       explained by one of the rules above: the flip propagated down the greedy chain.

Anything else is reported in ``unexplained`` and the callers fail on it.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def _iou_one_to_many(box, boxes):
    box = box.astype(np.float64)
    boxes = boxes.astype(np.float64)
    iw = np.minimum(box[2], boxes[:, 2]) - np.maximum(box[0], boxes[:, 0])
    ih = np.minimum(box[3], boxes[:, 3]) - np.maximum(box[1], boxes[:, 1])
    inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
    a = (box[2] - box[0]) * (box[3] - box[1])
    b = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (a + b - inter)


def explain_keep_mismatches(score: np.ndarray, xyxy: np.ndarray, n_per_image: int, keep_ref: np.ndarray,
                            keep_gpu: np.ndarray, iou_thr: float, score_thr: float, iou_band: float = 2e-6,
                            score_ulps: int = 4) -> Dict:
    """``score [B*N]``, ``xyxy [B*N,4]``: the oracle's fp32 scores and boxes of every candidate (flat index
    ``b*N + i``); ``keep_ref`` / ``keep_gpu``: the two keep-lists (flat indices, any order, after the score
    threshold)."""
    keep_ref = np.asarray(keep_ref, np.int64)
    keep_gpu = np.asarray(keep_gpu, np.int64)
    miss = np.setxor1d(keep_ref, keep_gpu)
    out = {"kept_ref": int(keep_ref.size), "kept_gpu": int(keep_gpu.size), "mismatches": int(miss.size),
           "score_threshold": 0, "iou_at_threshold": 0, "score_tie": 0, "cascade": 0, "unexplained": []}
    if miss.size == 0:
        return out
    union = np.union1d(keep_ref, keep_gpu)
    img_of_union = union // n_per_image
    explained = set()
    order = sorted(miss.tolist(), key=lambda c: (-float(score[c]), c))
    thr32 = np.float32(score_thr)
    for c in order:
        sc = np.float32(score[c])
        if abs(float(sc) - float(thr32)) <= score_ulps * float(np.spacing(max(abs(thr32), np.float32(1e-30)))):
            out["score_threshold"] += 1
            explained.add(c)
            continue
        others = union[img_of_union == c // n_per_image]
        others = others[others != c]
        if others.size == 0:
            out["unexplained"].append(int(c))
            continue
        iou = _iou_one_to_many(xyxy[c], xyxy[others])
        so = score[others].astype(np.float64)
        tie_eps = score_ulps * float(np.spacing(sc))
        ranked_above = so >= float(sc) - tie_eps          # could be processed before c by either implementation
        if np.any(ranked_above & (np.abs(iou - iou_thr) <= iou_band)):
            out["iou_at_threshold"] += 1
            explained.add(c)
            continue
        over = iou > iou_thr - iou_band
        if np.any(over & (np.abs(so - float(sc)) <= tie_eps)):
            out["score_tie"] += 1
            explained.add(c)
            continue
        if np.any(over & ranked_above & np.isin(others, np.fromiter(explained, np.int64, len(explained)))):
            out["cascade"] += 1
            explained.add(c)
            continue
        out["unexplained"].append(int(c))
    return out


def summarize(rep: Dict) -> str:
    return ("kept %d (oracle) / %d (cuda), mismatches %d [score_threshold %d, iou_at_threshold %d, score_tie %d, "
            "cascade %d, unexplained %d]" % (rep["kept_ref"], rep["kept_gpu"], rep["mismatches"], rep["score_threshold"],
                                             rep["iou_at_threshold"], rep["score_tie"], rep["cascade"],
                                             len(rep["unexplained"])))
