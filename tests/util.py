"""Shared helpers for the parity tests."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def digest(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t).tobytes())
    return h.hexdigest()[:16]


def canon(keep, scores):
    """(score desc, index asc) canonical order of a keep list."""
    keep = np.asarray(keep, dtype=np.int64)
    s = np.asarray(scores, dtype=np.float32)[keep].astype(np.float64)
    return keep[np.lexsort((keep, -s))]


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = ~((err <= tol) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{what}: {bad.sum()} / {a.size} outside rtol={rtol} atol={atol}; max err {np.nanmax(err):.3e}"


def rows_canon(rows, img):
    """Order [K,6] (score, cls, x1,y1,x2,y2) rows by image, score descending, then the remaining columns
    (the reference's order inside equal-score runs is arbitrary -- SURVEY A.4)."""
    rows = np.asarray(rows)
    if rows.shape[0] == 0:
        return rows
    r = rows.astype(np.float64)
    order = np.lexsort((r[:, 5], r[:, 4], r[:, 3], r[:, 2], r[:, 1], -r[:, 0], np.asarray(img)))
    return rows[order]


def assign_variant_case(name):
    """(targets, overlap_masks, batch_size) of a tests/golden/assign_variants.npz case (oracle/make_golden.py)."""
    from vision_conglomerate_b200 import synth
    return {
        "seg_overlap": lambda: (synth.targets(4, 12, 80, 3, fixed=False), True, 4),
        "seg_plain": lambda: (synth.targets(4, 12, 80, 3, fixed=False), False, None),
        "kpt": lambda: (synth.keypoint_targets(3, 10, 2), None, None),
        "kpt_seg_overlap": lambda: (synth.keypoint_targets(3, 10, 2), True, 3),
    }[name]()


ASSIGN_VARIANTS = ("seg_overlap", "seg_plain", "kpt", "kpt_seg_overlap")


def seg_extra_columns(B, N, n_extra, seed):
    """The seeded mask-coefficient columns oracle/make_golden.py appended to the decoded rows (gen_seg_post)."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    return torch.tanh(torch.randn(B, N, n_extra, generator=g))


def rows_order(rows, img):
    """The permutation rows_canon applies."""
    r = np.asarray(rows).astype(np.float64)
    if r.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    return np.lexsort((r[:, 5], r[:, 4], r[:, 3], r[:, 2], r[:, 1], -r[:, 0], np.asarray(img)))


def seg_loss_case(name, g):
    """Inputs of a tests/golden/segloss_*.npz case (oracle/make_golden.py gen_seg_loss): preds, protos, targets, masks, C, K."""
    from vision_conglomerate_b200 import synth
    B, H, W, C, K, G, seed, mdiv, fixed = (int(v) for v in g["params"])
    preds, protos, t, masks = synth.seg_inputs(B, H, W, C, K, G, seed, mdiv, bool(fixed))
    if name == "segloss_empty_img":
        t = t[t[:, 0] != 1].contiguous()
    return preds, protos, t, masks, C, K


def seg_mask_case(g):
    """Inputs of a tests/golden/segmask_*.npz case (oracle/make_golden.py gen_seg_masks): raw head outputs, decode
    parameters, the K = 4 extra columns' seed, the protos, the image size and the reference's per-image rows / masks."""
    import torch
    from vision_conglomerate_b200 import synth
    gd = golden(str(g["decode_case"]))
    B, H, W, C, seed, og0, og1 = (int(v) for v in gd["params"])
    raws = synth.raw_head_outputs(B, H, W, C, str(gd["dist"]), seed)
    psz, isz = [int(v) for v in g["psz"]], [int(v) for v in g["isz"]]
    protos = torch.randn(B, 4, psz[0], psz[1], generator=torch.Generator().manual_seed(int(g["proto_seed"]))).contiguous()
    nrow = int(g["mask_rows"])
    masks = np.unpackbits(g["masks_packed"], axis=1)[:, : isz[0] * isz[1]].reshape(nrow, isz[0], isz[1]).astype(bool)
    return raws, (B, H, W, C, None if og0 < 0 else (og0, og1)), protos, isz, masks


def assert_masks_match(got, ref, vals, what="masks", band=2e-5):
    """Boolean masks must agree except where the interpolated value sits on the 0.5 threshold to within `band`
    (cuBLAS / ATen and the restatement add the K products and the four bilinear terms in different orders)."""
    got, ref = np.asarray(got, bool), np.asarray(ref, bool)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    diff = got != ref
    marginal = np.abs(np.asarray(vals, np.float64) - 0.5) < band
    assert not (diff & ~marginal).any(), f"{what}: {int((diff & ~marginal).sum())} pixels differ away from the threshold"
    return int(diff.sum())
