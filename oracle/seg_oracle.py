"""CPU restatement (numpy, fp32) of the mask term of the reference's ``SegmentationLoss``
(modules/segmentation_loss.py:26-231, ``overlap_masks=True``, BCEWithLogits) and of the whole ``forward`` on top of the
C oracle's detection terms.  TEST INFRASTRUCTURE: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
may import this; pinned against the unmodified reference by ``tests/golden/segloss_*.npz`` (``oracle/make_golden.py``).

Per scale and image i with matches m (segmentation_loss.py:155-171, 208-231; utils/utils.py:130-172):

    pred_m   = coefs_m @ protos_i                      [Hp*Wp]      (:220; coefs = preds[b,gj,gi,a, 5+C : 5+C+K])
    t_m      = (target_masks_i == tmask_idx_m)         0/1          (:162; masks nearest-resized to the protos' size, :152-153)
    bce      = BCEWithLogits(pred_m, t_m)              per pixel    (:222)
    dice_i   = mean_m (2 sum(sig*t) + e) / (sum(sig) + sum(t) + e)        e = 1e-5   (:223, utils.py:169-171)
    L_m      = sum(bce * crop_m) / (Hp*Wp) / (w_m*h_m)              (:225; crop_m = pixels r,c with x1<=r<x2, y1<=c<y2 of the
                                                                     match's own (x,y,w,h) -- grid units, as the reference does)
    sl_i     = mean_m (1 - L_m) * (1 - dice_i)                      (:226-227)
    seg_loss = sum_i sl_i / B ;  dice_score = sum_i mean_m dice(round(sig), t) / B   (:168-171, :224)
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import oracle as O

DICE_E = np.float32(1e-5)


def _sigmoid(x):
    x = x.astype(np.float32)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = np.float32(1) / (np.float32(1) + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (np.float32(1) + ex)
    return out


def _bce_logits(x, t):
    # ATen: (1 - t) * x - log_sigmoid(x),  log_sigmoid(x) = min(x, 0) - log1p(exp(-|x|))
    return ((np.float32(1) - t) * x - (np.minimum(x, np.float32(0)) - np.log1p(np.exp(-np.abs(x))))).astype(np.float32)


def resize_nearest(masks: np.ndarray, Hp: int, Wp: int) -> np.ndarray:
    """``F.interpolate(mode="nearest")`` (:152-153): src = min(floor(dst * in/out), in-1), the scale in fp32."""
    B, Hm, Wm = masks.shape
    if (Hm, Wm) == (Hp, Wp):
        return masks
    sy, sx = np.float32(Hm) / np.float32(Hp), np.float32(Wm) / np.float32(Wp)
    iy = np.minimum(np.floor(np.arange(Hp, dtype=np.float32) * sy).astype(np.int64), Hm - 1)
    ix = np.minimum(np.floor(np.arange(Wp, dtype=np.float32) * sx).astype(np.int64), Wm - 1)
    return masks[:, iy][:, :, ix]


def mask_term_scale(preds, targets, anchors, protos, target_masks, C: int, K: int, cfg: dict, with_grad: bool = False):
    """One scale of the mask term.  Returns (seg_loss, dice_score, grad_coef_rows or None, grad_protos or None):
    ``grad_coef_rows`` is dense like ``preds`` (non-zero only in the K coefficient columns of matched rows)."""
    preds = np.asarray(preds, np.float32)
    protos = np.asarray(protos, np.float32)
    B, ny, nx, na, D = preds.shape
    _, Kp, Hp, Wp = protos.shape
    assert Kp == K and D >= 5 + C + K
    idx, cls, anc, box, tmask, _ = O.build_target_by_scale_ex(targets, (ny, nx), anchors, cfg["anchor_t"], cfg["edge_t"],
                                                              overlap_masks=True, batch_size=B)
    tm = resize_nearest(np.asarray(target_masks, np.float32), Hp, Wp)
    bi, gj, gi, ai = idx
    seg_loss, dice_score = np.float32(0), np.float32(0)
    g_rows = np.zeros_like(preds) if with_grad else None
    g_protos = np.zeros_like(protos) if with_grad else None
    r = np.arange(Wp, dtype=np.float32)[None, None, :]
    c = np.arange(Hp, dtype=np.float32)[None, :, None]
    for i in np.unique(bi):
        m = np.nonzero(bi == i)[0]
        n = m.shape[0]
        coefs = preds[bi[m], gj[m], gi[m], ai[m], 5 + C:5 + C + K]                 # [n, K]
        P = protos[i].reshape(K, -1)
        pred = (coefs @ P).reshape(n, Hp, Wp).astype(np.float32)
        t = (tm[i][None] == tmask[m].astype(np.float32).reshape(-1, 1, 1)).astype(np.float32)
        sig = _sigmoid(pred)
        bce = _bce_logits(pred, t)
        inter = (sig * t).sum(axis=(1, 2), dtype=np.float32)
        den = sig.sum(axis=(1, 2), dtype=np.float32) + t.sum(axis=(1, 2), dtype=np.float32)
        dice = (np.float32(2) * inter + DICE_E) / (den + DICE_E)
        dice_loss = np.float32(1) - dice.mean(dtype=np.float32)
        sr = np.round(sig)                                  # half to even, like torch.round
        inter_r = (sr * t).sum(axis=(1, 2), dtype=np.float32)
        den_r = sr.sum(axis=(1, 2), dtype=np.float32) + t.sum(axis=(1, 2), dtype=np.float32)
        ds = ((np.float32(2) * inter_r + DICE_E) / (den_r + DICE_E)).mean(dtype=np.float32)
        b4 = box[m]
        x1 = (b4[:, 0] - b4[:, 2] / np.float32(2))[:, None, None]
        y1 = (b4[:, 1] - b4[:, 3] / np.float32(2))[:, None, None]
        x2 = (b4[:, 0] + b4[:, 2] / np.float32(2))[:, None, None]
        y2 = (b4[:, 1] + b4[:, 3] / np.float32(2))[:, None, None]
        crop = ((r >= x1) & (r < x2) & (c >= y1) & (c < y2)).astype(np.float32)
        area = (b4[:, 2] * b4[:, 3]).astype(np.float32)
        Lm = (bce * crop).mean(axis=(1, 2), dtype=np.float32) / area
        A = (np.float32(1) - Lm).mean(dtype=np.float32)
        seg_loss += A * dice_loss
        dice_score += ds
        if with_grad:
            # d sl_i / d pred[m, px], sl_i = A * dice_loss
            dL = -dice_loss / np.float32(n)                 # d sl / d L_m
            dD = -A / np.float32(n)                         # d sl / d dice_m
            Den = (den + DICE_E)[:, None, None]
            Num = (np.float32(2) * inter + DICE_E)[:, None, None]
            dpred = dL * crop * (sig - t) / np.float32(Hp * Wp) / area[:, None, None] \
                + dD * sig * (np.float32(1) - sig) * (np.float32(2) * t * Den - Num) / (Den * Den)
            dpred = (dpred / np.float32(B)).astype(np.float32).reshape(n, -1)
            gc = dpred @ P.T                                # [n, K]
            np.add.at(g_rows, (bi[m], gj[m], gi[m], ai[m]), np.concatenate(
                [np.zeros((n, 5 + C), np.float32), gc.astype(np.float32), np.zeros((n, D - 5 - C - K), np.float32)], 1))
            g_protos[i] += (coefs.T @ dpred).reshape(K, Hp, Wp)
    return float(seg_loss / np.float32(B)), float(dice_score / np.float32(B)), g_rows, g_protos


def segmentation_loss(preds3: Sequence, targets, protos, target_masks, anchors3: Sequence, cfg: dict, C: int, K: int,
                      with_grad: bool = False):
    """``SegmentationLoss.forward`` (modules/segmentation_loss.py:26-75): detection terms from the C oracle on the
    first 5+C columns, the mask term from above.  Returns (loss, metrics, [grad_sm, grad_md, grad_lg] or None,
    grad_protos or None)."""
    preds3 = [np.asarray(p, np.float32) for p in preds3]
    det3 = [np.ascontiguousarray(p[..., :5 + C]) for p in preds3]
    loss, metrics, dgr, _ = O.detection_loss(det3, targets, anchors3, cfg, with_grad)
    sw = cfg.get("scale_w") or [4.0, 2.0, 1.0]
    lseg, seg_rows, dice_rows, grads, gp = 0.0, [], [], [], None
    for s, (p, a, w) in enumerate(zip(preds3, anchors3, sw)):
        sl, ds, gr, gpr = mask_term_scale(p, targets, a, protos, target_masks, C, K, cfg, with_grad)
        lseg += w * sl
        seg_rows.append(sl)
        dice_rows.append(ds)
        if with_grad:
            k = np.float32(cfg.get("seg_w", 1.0) * w)
            g = gr * k
            g[..., :5 + C] += dgr[s]
            grads.append(g)
            gp = gpr * k if gp is None else gp + gpr * k
    loss = loss + cfg.get("seg_w", 1.0) * lseg
    metrics = dict(metrics, aggregate_loss=float(loss), seg_loss=float(np.mean(seg_rows)), dice_score=float(np.mean(dice_rows)))
    return float(loss), metrics, (grads if with_grad else None), gp


def seg_masks(coefs, row_counts, protos, H: int, W: int):
    """inference_seg.post_process_preds lines 115-117: per image, ``masks = sigmoid(coefs @ protos_i)`` on the protos'
    grid, ``F.interpolate(mode="bilinear", align_corners=False)`` to ``(H, W)``, ``> 0.5``.  ``coefs [n, K]`` are the kept
    rows' coefficients image by image (``row_counts [B]``).  Returns (bool [n, H, W], the interpolated values fp32).
    ATen's bilinear kernel: src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out in fp32, i0 = trunc(src),
    i1 = i0 + (i0 < in - 1), value = h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)."""
    coefs = np.asarray(coefs, np.float32)
    protos = np.asarray(protos, np.float32)
    B, K, Hp, Wp = protos.shape
    n = coefs.shape[0]
    vals = np.zeros((n, H, W), np.float32)

    def axis(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = np.maximum(scale * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5), np.float32(0))
        i0 = src.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (np.float32(1) - l1).astype(np.float32), l1

    y0, y1, hy0, hy1 = axis(Hp, H)
    x0, x1, wx0, wx1 = axis(Wp, W)
    r = 0
    for i, c in enumerate(np.asarray(row_counts, np.int64)):
        if c == 0:
            continue
        low = _sigmoid((coefs[r:r + c] @ protos[i].reshape(K, -1)).astype(np.float32)).reshape(c, Hp, Wp)
        top = wx0[None, None, :] * low[:, y0][:, :, x0] + wx1[None, None, :] * low[:, y0][:, :, x1]
        bot = wx0[None, None, :] * low[:, y1][:, :, x0] + wx1[None, None, :] * low[:, y1][:, :, x1]
        vals[r:r + c] = hy0[None, :, None] * top + hy1[None, :, None] * bot
        r += c
    return vals > np.float32(0.5), vals
