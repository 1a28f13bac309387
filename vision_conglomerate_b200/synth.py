"""Seeded synthetic inputs for the box-geometry hot path.

No dataset or checkpoint is available offline, so every test and benchmark draws
its inputs from here.  Everything is generated on the CPU with explicit
``torch.Generator`` seeds so the oracle, the golden fixtures and the CUDA path
all see bit-identical tensors.

Distributions of raw head outputs ``[B, S/s, S/s, 3, 5+C]`` (channels
``[obj, cls*C, tx, ty, tw, th]``, the layout produced by the reference's
``EffiDecHead.forward``, modules/common.py:912-931):

* ``R``  random-init like: every channel ``0.0098 + 0.0043*n`` -- every candidate
  clears conf 0.001 (NMS worst case, ALU bound).
* ``T``  trained like: obj ``-9 + 2n``, cls ``-4 + 1.5n``, box ``n`` -- about 6.8 %
  of candidates clear conf 0.001.
* ``TP`` ``T`` plus 64 planted confident candidates per image (score about 0.96)
  so a 0.3 score threshold leaves a non-empty result.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

# config/detection/anchors.yaml of the reference, verbatim values (w, h) normalised to the image.
ANCHORS: Dict[str, List[List[float]]] = {
    "sm": [[0.03421874716877937, 0.11828124523162842],
           [0.04921875149011612, 0.09843750298023224],
           [0.05546874925494194, 0.09999999403953552]],
    "md": [[0.05898437649011612, 0.20078125596046448],
           [0.06562499701976776, 0.3382812440395355],
           [0.08281250298023224, 0.34687501192092896]],
    "lg": [[0.09375, 0.34687501192092896],
           [0.1171875, 0.2945312559604645],
           [0.10703124850988388, 0.3359375]],
}
STRIDES = (8, 16, 32)
SCALES = ("sm", "md", "lg")

# config/detection/config.yaml:65-76 loss_config of the reference.
LOSS_CONFIG = dict(alpha=None, anchor_t=4.0, batch_scale_loss=False, box_w=0.1, class_w=0.3,
                   conf_w=1.0, keypoints_w=5.0, edge_t=0.5, gamma=None, label_smoothing=0.001,
                   scale_w=[4.0, 2.0, 1.0])


def anchors_tensor(scale: str) -> torch.Tensor:
    return torch.tensor(ANCHORS[scale], dtype=torch.float32)


def fmap_shapes(H: int, W: int) -> List[Tuple[int, int]]:
    return [(H // s, W // s) for s in STRIDES]


def raw_head_outputs(B: int, H: int, W: int, C: int = 80, dist: str = "T", seed: int = 7,
                     na: int = 3, planted: int = 64) -> List[torch.Tensor]:
    """Three raw head tensors ``[B, ny, nx, na, 5+C]`` fp32, drawn scale by scale."""
    g = torch.Generator().manual_seed(seed)
    D = 5 + C
    outs = []
    for ny, nx in fmap_shapes(H, W):
        n = torch.randn(B, ny, nx, na, D, generator=g, dtype=torch.float32)
        if dist == "R":
            z = 0.0098 + 0.0043 * n
        elif dist in ("T", "TP"):
            z = n.clone()
            z[..., 0] = -9.0 + 2.0 * n[..., 0]
            z[..., 1:1 + C] = -4.0 + 1.5 * n[..., 1:1 + C]
        elif dist == "N":  # plain standard normal (used for small parity cases)
            z = n
        else:
            raise ValueError(dist)
        outs.append(z.contiguous())
    if dist == "TP":
        sizes = [o.shape[1] * o.shape[2] * na for o in outs]
        N = sum(sizes)
        for b in range(B):
            gp = torch.Generator().manual_seed(1000 + seed - 7 + b)
            flat = torch.randint(N, (planted,), generator=gp)
            cls = torch.randint(C, (planted,), generator=gp)
            for f, c in zip(flat.tolist(), cls.tolist()):
                s = 0
                while f >= sizes[s]:
                    f -= sizes[s]
                    s += 1
                row = outs[s][b].reshape(-1, D)[f]
                row[0] = 4.0
                row[1 + c] = 4.0
    return outs


def targets(B: int, G: int, C: int = 80, seed: int = 0, fixed: bool = True) -> torch.Tensor:
    """``[nt, 6]`` fp32 rows ``(img, cls, x, y, w, h)`` normalised, concatenated in image order
    (what the reference's ``DetectionDataset.collate_fn`` produces, detection_dataset.py:81-88)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(B):
        n = G if fixed else int(torch.randint(1, G + 1, (1,), generator=g))
        cls = torch.randint(0, C, (n,), generator=g).float()
        xy = 0.05 + 0.9 * torch.rand(n, 2, generator=g)
        wh = 0.02 + 0.3 * torch.rand(n, 2, generator=g)
        rows.append(torch.cat([torch.full((n, 1), float(b)), cls[:, None], xy, wh], dim=1))
    return torch.cat(rows, 0).contiguous() if rows else torch.zeros(0, 6)


def adversarial_targets(B: int = 2, C: int = 80) -> torch.Tensor:
    """Border / cell-boundary / duplicate targets that exercise the clamp and the edge rules."""
    vals = [0.0, 1.0, 0.5, 0.25, 1.0 / 80, 1.5 / 80, 0.999999, 1e-6, 0.0125, 0.9875, 0.0126, 79.5 / 80]
    rows = []
    for b in range(B):
        for i, x in enumerate(vals):
            for j, y in enumerate(vals[: 6]):
                rows.append([b, (i * 7 + j) % C, x, y, 0.05 + 0.01 * j, 0.12 + 0.02 * i])
        rows.append([b, 3, 0.4, 0.4, 0.06, 0.2])
        rows.append([b, 3, 0.4, 0.4, 0.06, 0.2])      # exact duplicate
        rows.append([b, 5, 0.4, 0.4, 0.9, 0.9])       # fails every anchor ratio
        rows.append([b, 5, 0.4, 0.4, 0.001, 0.001])   # fails every anchor ratio (too small)
    return torch.tensor(rows, dtype=torch.float32)


def keypoint_targets(B: int, G: int, num_kp: int = 2, C: int = 80, seed: int = 5) -> torch.Tensor:
    """``targets`` with ``3 * num_kp`` keypoint columns (x, y, visibility) appended, variable boxes per image."""
    t = targets(B, G, C, seed, fixed=False)
    g = torch.Generator().manual_seed(seed + 100)
    kp = torch.rand(t.shape[0], num_kp, 3, generator=g)
    kp[..., 2] = torch.randint(0, 3, (t.shape[0], num_kp), generator=g).float()
    return torch.cat([t, kp.reshape(t.shape[0], -1)], 1).contiguous()


def train_preds(B: int, H: int, W: int, C: int = 80, seed: int = 1, na: int = 3) -> List[torch.Tensor]:
    """Training-mode (already decoded) prediction tensors for the loss, ``torch.manual_seed(seed)``."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, ny, nx, na, 5 + C, generator=g, dtype=torch.float32) for ny, nx in fmap_shapes(H, W)]


def seg_inputs(B: int, H: int, W: int, C: int, K: int, G: int, seed: int = 3, mask_div: int = 1, fixed: bool = False):
    """Seeded inputs of ``SegmentationLoss.forward`` (modules/segmentation_loss.py:26-75) with ``overlap_masks=True``:
    training-mode prediction tensors with ``K`` mask-coefficient columns (``tanh``-ranged, detection.py:131-134),
    ``protos [B, K, H/2, W/2]`` (modules/segmentation.py:21), targets, and the overlapped target masks
    ``[B, H/mask_div, W/mask_div]`` whose pixels hold 1 + the position of the covering object inside its image
    (later objects drawn over earlier ones; utils/utils.py polygons_2_overlapped_mask), here the objects' boxes."""
    g = torch.Generator().manual_seed(seed)
    preds = []
    for ny, nx in fmap_shapes(H, W):
        p = torch.randn(B, ny, nx, 3, 5 + C + K, generator=g, dtype=torch.float32)
        p[..., 5 + C:] = torch.tanh(p[..., 5 + C:])
        preds.append(p.contiguous())
    protos = torch.randn(B, K, H // 2, W // 2, generator=g, dtype=torch.float32)
    t = targets(B, G, C, seed + 1, fixed)
    Hm, Wm = H // mask_div, W // mask_div
    masks = torch.zeros(B, Hm, Wm, dtype=torch.float32)
    pos = {}
    for row in t.tolist():
        b = int(row[0])
        pos[b] = pos.get(b, 0) + 1
        x, y, w, h = row[2:6]
        x1, x2 = int(max(0.0, x - w / 2) * Wm), int(min(1.0, x + w / 2) * Wm) + 1
        y1, y2 = int(max(0.0, y - h / 2) * Hm), int(min(1.0, y + h / 2) * Hm) + 1
        masks[b, y1:y2, x1:x2] = float(pos[b])
    return preds, protos, t, masks


def nms_boxes(n: int, groups: int, seed: int = 3, extent: float = 640.0, ties: bool = False
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Random xyxy boxes, scores and int64 group ids for stand-alone NMS tests."""
    g = torch.Generator().manual_seed(seed)
    cxy = torch.rand(n, 2, generator=g) * extent
    wh = 8.0 + torch.rand(n, 2, generator=g) * extent * 0.25
    boxes = torch.cat([cxy - wh / 2, cxy + wh / 2], 1).contiguous()
    scores = torch.rand(n, generator=g)
    if ties:
        scores = (scores * 16).floor() / 16  # heavy score ties
        boxes[n // 2:] = boxes[: n - n // 2].clone()   # exact duplicate boxes
    idxs = torch.randint(0, groups, (n,), generator=g, dtype=torch.int64)
    return boxes, scores.contiguous(), idxs


def candidates_per_image(H: int, W: int, na: int = 3) -> int:
    return sum(ny * nx * na for ny, nx in fmap_shapes(H, W))


def tracked_classes_default() -> Sequence[int]:
    return (1, 4, 7, 16, 17)
