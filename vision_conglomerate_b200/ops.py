"""Torch-facing operators over libboxgeom.so.

PyTorch is plumbing here: it owns device memory (inputs, outputs, workspaces come from the caching
allocator) and supplies the CUDA stream; all arithmetic happens in the hand-written sm_100a kernels.
Every operator rejects CPU / non-fp32 tensors -- there is no fallback path.

Each function cites the reference routine whose call signature it mirrors (paths relative to the
reference root).
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DetectParams, LossParams, check

_ws: Dict[Tuple[int, str], torch.Tensor] = {}
_pinned: Dict[Tuple[int, str], torch.Tensor] = {}
_mask_budget: Dict[Tuple[int, str], int] = {}
DEFAULT_MASK_BYTES = 64 << 20
# bg_detect NMS path remembered per (device, batch, candidates, thresholds): absent = per-image kernel for up to
# 4,096 survivors (default), 4 = per-image kernel for up to 8,192, 1 = general segmented engine (more survivors),
# 3 = general for good (overlap-edge overflow)
_nms_path_hint: Dict[Tuple, int] = {}
PER_IMAGE_NMS_CAP = 4096        # InmsSmall::CAP of csrc/imgnms_kernels.cuh
PER_IMAGE_NMS_CAP_LARGE = 8192  # InmsLarge::CAP
PER_IMAGE_NMS_CAP_LEAN = 2048   # InmsLean::CAP (throughput mode: a CTA that runs next to the decode CTAs of other streams)
LEAN_NMS = True                 # throughput plans start on the lean kernel (False: the 1024-thread kernel, one CTA per image)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream(dev: Optional[torch.device] = None) -> int:
    """cudaStream_t of torch's current stream on `dev` (default: the current device); the raw accessors skip
    ~10 us of Python."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device() if dev is None or dev.index is None else dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


def _on(dev: torch.device):
    """Context that makes `dev` the current CUDA device.  libboxgeom launches on the current device (its
    per-device tables, the stream handle and the pointers must all belong to it), while callers such as the
    reference's DDP path address ranks as ``cuda:k`` without ever calling ``torch.cuda.set_device``."""
    return torch.cuda.device(dev)


def _same_device(*tensors) -> torch.device:
    dev = tensors[0].device
    for t in tensors[1:]:
        if t is not None and t.device != dev:
            raise RuntimeError(f"all tensors of one call must live on one device (got {dev} and {t.device})")
    return dev


def _workspace(dev: torch.device, kind: str, nbytes: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), kind)
    t = _ws.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _ws[key] = t
    return t


def _pinned_i32(dev: torch.device, kind: str, n: int) -> torch.Tensor:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), kind)
    t = _pinned.get(key)
    if t is None or t.numel() < n:
        t = torch.empty(int(n), dtype=torch.int32).pin_memory()
        _pinned[key] = t
    return t


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: the box-geometry kernels need a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


_anchor_cache: Dict[int, tuple] = {}


def _anchors_host(anchors) -> List[List[float]]:
    """Anchor (w, h) pairs as python floats.  The model keeps them as device parameters (read at call time, like the
    reference does): the copy to the host -- a stream sync -- is remembered per tensor object and version."""
    if isinstance(anchors, torch.Tensor):
        ent = _anchor_cache.get(id(anchors))
        if ent is not None and ent[0]() is anchors and ent[1] == anchors._version:
            return ent[2]
        if len(_anchor_cache) > 256:
            _anchor_cache.clear()
        val = anchors.detach().to("cpu", torch.float32).reshape(-1, 2).tolist()
        _anchor_cache[id(anchors)] = (weakref.ref(anchors), anchors._version, val)
        return val
    return torch.tensor(anchors, dtype=torch.float32).reshape(-1, 2).tolist()


def _anchor_array(anchors) -> "C.Array":
    a = _anchors_host(anchors)
    arr = (C.c_float * (2 * len(a)))()
    for i, (w, h) in enumerate(a):
        arr[2 * i] = w
        arr[2 * i + 1] = h
    return arr


def _read_counts(dev_counts: torch.Tensor, kind: str) -> torch.Tensor:
    """One device->host copy + one stream sync: the only host-visible sync of an operator."""
    host = _pinned_i32(dev_counts.device, kind, dev_counts.numel())[: dev_counts.numel()]
    host.copy_(dev_counts, non_blocking=True)
    torch.cuda.current_stream(dev_counts.device).synchronize()
    return host


# ---------------------------------------------------------------------------------------------- B4
def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                max_groups: int = 1 << 16) -> torch.Tensor:
    """Drop-in for ``torchvision.ops.batched_nms`` as called at inference_det.py:77-82: greedy NMS per
    distinct ``idxs`` value, int64 keep indices, score-descending (index-ascending inside equal scores)."""
    boxes = _req(boxes, "boxes").reshape(-1, 4)
    scores = _req(scores, "scores").reshape(-1)
    idxs = _req(idxs, "idxs", torch.int64).reshape(-1)
    n = scores.numel()
    dev = _same_device(boxes, scores, idxs)
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    with _on(dev):
        return _batched_nms_on(dev, boxes, scores, idxs, n, iou_threshold, max_groups)


def _batched_nms_on(dev, boxes, scores, idxs, n, iou_threshold, max_groups):
    L = _lib.lib()
    keep = torch.empty(n, dtype=torch.int64, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    key = (dev.index, "gnms")
    mask_bytes = _mask_budget.get(key, DEFAULT_MASK_BYTES)
    while True:
        ws = _workspace(dev, "gnms", L.bg_batched_nms_workspace_bytes(n, max_groups, mask_bytes))
        check(L.bg_batched_nms(boxes.data_ptr(), scores.data_ptr(), idxs.data_ptr(), n, float(iou_threshold),
                               max_groups, keep.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(), mask_bytes,
                               _stream(dev)), "bg_batched_nms")
        h = _read_counts(counts, "gnms")
        status = int(h[1])
        if status & _lib.STATUS_GROUP_RANGE:
            if max_groups >= (1 << 24):
                raise RuntimeError("batched_nms: idxs span more than 2^24 distinct values")
            max_groups = min(max_groups << 4, 1 << 24)
            continue
        if status & _lib.STATUS_MASK_SPACE:
            worst = n * ((n + 63) // 64) * 8
            if mask_bytes >= worst:
                raise RuntimeError("batched_nms: suppression-mask workspace exhausted")
            mask_bytes = min(worst, mask_bytes * 8)
            _mask_budget[key] = mask_bytes
            continue
        return keep[: int(h[0])]


# ---------------------------------------------------------------------------------------------- B5
@dataclass
class Detections:
    """Result of :func:`detect`.  ``pred_boxes[:, :] = (score, class, x1, y1, x2, y2)`` (inference_det.py:93-97),
    ``sample_idxs`` the image of each row (:87), ``keep_idxs`` its flat candidate index ``b*N + i``,
    ``counts[b]`` rows of image b (rows are image-major unless ``order='global'``), ``candidates[b]`` how many
    candidates of image b cleared the score threshold."""
    pred_boxes: torch.Tensor
    sample_idxs: torch.Tensor
    keep_idxs: torch.Tensor
    counts: torch.Tensor
    candidates: torch.Tensor


class DetectPlan:
    """Parameters, output buffers and workspace of one fused decode+NMS configuration, reusable across
    batches of the same shape.  ``enqueue`` launches the kernels on the current stream without any host
    synchronisation; ``result`` performs the single device->host read the reference signature forces.
    ``throughput=True`` tunes the launches for several batches in flight (see :class:`DetectPipeline`): one NMS CTA
    per image instead of main + helper, no programmatic dependent launch.
    ``host_result=True`` (batch-1 video frames, BASELINE config 5): the output buffers live in page-locked host memory
    that the kernels write directly, and the last kernel stores a sequence number there (``bg_detect_params.host_flag``);
    ``result_host`` then polls that word -- no device->host copy, no stream synchronisation.  Meant for outputs of a few
    hundred rows: every row crosses PCIe as the kernel writes it."""

    def __init__(self, shapes: Sequence[Tuple[int, ...]], anchors3: Sequence, input_shape: Tuple[int, int],
                 num_classes: int, device: torch.device, og_size: Optional[Tuple[int, int]] = None,
                 iou_threshold: float = 0.5, score_threshold: float = 0.1, box_allowance: Optional[float] = None,
                 tracked_classes: Optional[Sequence[int]] = None, order: str = "image", variant: int = 0,
                 nms_path: str = "auto", predecoded: bool = False, throughput: bool = False, host_result: bool = False):
        self.predecoded = bool(predecoded)
        self.host_result = bool(host_result)
        if len(shapes) != 3 or any(len(sh) != 5 for sh in shapes):
            raise RuntimeError("detect: expected three [B, ny, nx, na, 5+C] head outputs")
        B, _, _, na, D = shapes[0]
        if D < num_classes + 5:
            raise RuntimeError("detect: rows must hold at least 5 + num_classes columns")
        p = DetectParams()
        p.B, p.C, p.na = B, num_classes, na
        # trailing columns (mask coefficients / keypoints of the segmentation and keypoint heads, inference_seg.py:66-68)
        # take no part in the box geometry: the kernels skip them, `extra_columns` gathers them for the kept rows
        p.extra_cols = D - (num_classes + 5)
        p.H, p.W = int(input_shape[0]), int(input_shape[1])
        p.og_H, p.og_W = (int(og_size[0]), int(og_size[1])) if og_size is not None else (-1, -1)
        for s, sh in enumerate(shapes):
            if sh[0] != B or sh[3] != na or sh[4] != D:
                raise RuntimeError("detect: inconsistent head output shapes")
            p.ny[s], p.nx[s] = sh[1], sh[2]
            for a, (w, h) in enumerate(_anchors_host(anchors3[s])):
                p.anchors[s][a][0] = w
                p.anchors[s][a][1] = h
        p.box_allowance = float(box_allowance) if box_allowance else 0.0
        p.score_threshold = float(score_threshold)
        p.iou_threshold = float(iou_threshold)
        tracked = list(tracked_classes) if tracked_classes else []
        if len(tracked) > _lib.BG_MAX_TRACKED:
            raise RuntimeError("detect: at most %d tracked classes" % _lib.BG_MAX_TRACKED)
        p.n_tracked = len(tracked)
        for i, c in enumerate(tracked):
            p.tracked[i] = int(c)
        p.order = 1 if order == "global" else 0
        p.variant = int(variant)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.params, self.B, self.dev = p, B, device
        self.N = sum(sh[1] * sh[2] * na for sh in shapes)
        self.hint_key = (device.index, B, self.N, p.iou_threshold, p.score_threshold)
        hint = _nms_path_hint.get(self.hint_key, 0)
        if nms_path == "general":
            p.nms_path = 1
        elif nms_path in ("per_image", "per_image_single", "per_image_large", "per_image_lean"):
            p.nms_path = {"per_image": 2, "per_image_single": 3, "per_image_large": 4, "per_image_lean": 5}[nms_path]
        else:
            p.nms_path = self._auto_path(hint, bool(throughput))
        p.throughput = 1 if throughput else 0
        self.shapes = [tuple(sh) for sh in shapes]
        n = B * self.N
        # counts and rows share one device allocation ([counts | pad | rows]) so that result_host() can bring both to the
        # host with ONE copy
        self._hdr_bytes = ((3 + 2 * B) * 4 + 255) // 256 * 256        # (one more word: the host flag of host_result plans)
        self._host_packed = None
        self._rows_guess = 256
        if self.host_result:
            # [counts | flag | pad | rows], the image and candidate index of every row: page-locked host memory, written by
            # the kernels through the device's mapping of it (under unified addressing the same address)
            self._packed = torch.empty(self._hdr_bytes + n * 24, dtype=torch.uint8).pin_memory()
            self.out_img = torch.empty(n, dtype=torch.int64).pin_memory()
            self.out_keep = torch.empty(n, dtype=torch.int64).pin_memory()
            with _on(device):
                for t in (self._packed, self.out_img, self.out_keep):
                    if _lib.lib().bg_host_mapped_ptr(t.data_ptr()) != t.data_ptr():
                        raise RuntimeError("detect: page-locked host memory is not addressable by the device at its own address")
            self._np_hdr = self._packed[: self._hdr_bytes].view(torch.int32).numpy()
            self._np_rows = self._packed[self._hdr_bytes:].view(torch.float32).view(n, 6).numpy()
            self._flag_idx, self._seq = 2 + 2 * B, 0
            self._np_hdr[self._flag_idx] = 0
            p.host_flag = self._packed.data_ptr() + 4 * self._flag_idx
        else:
            self._packed = torch.empty(self._hdr_bytes + n * 24, dtype=torch.uint8, device=device)
            self.out_img = torch.empty(n, dtype=torch.int64, device=device)
            self.out_keep = torch.empty(n, dtype=torch.int64, device=device)
        self.out_boxes = self._packed[self._hdr_bytes:].view(torch.float32).view(n, 6)
        self.counts = self._packed[: (2 + 2 * B) * 4].view(torch.int32)
        self.key = (device.index, "detect")
        self.ws_tag = "detect"   # plans that run concurrently on different streams need distinct scratch: set a distinct tag
        self.input_bytes = sum(4 * B * sh[1] * sh[2] * na * D for sh in shapes)
        self._need_key, self._need = None, 0
        self._out_ptrs = (self.out_boxes.data_ptr(), self.out_img.data_ptr(), self.out_keep.data_ptr(), self.counts.data_ptr())

    def _auto_path(self, hint: int, throughput: bool) -> int:
        """``bg_detect_params.nms_path`` for what earlier batches of this configuration needed (``_nms_path_hint``): 0 nothing
        known / sparse, 2 more than the lean kernel holds, 4 more than the 4 096-survivor kernel holds, else general."""
        lean_ok = throughput and LEAN_NMS   # (the library steps to the 1024-thread kernel when an image has >= 512 tiles)
        if hint == 0:
            return 5 if lean_ok else 0
        if hint == 2:
            return 2 if throughput else 0
        return 4 if hint == 4 else 1

    def use_nms_stream(self, stream: Optional["torch.cuda.Stream"]) -> None:
        """Run the NMS kernels on ``stream`` (normally one of higher priority than the stream ``enqueue`` is called on),
        ordered behind the decode kernel and in front of whatever follows on the enqueueing stream by an event of the
        plan (``bg_detect_params.nms_stream`` / ``nms_event``).  ``None`` switches it off."""
        if stream is None:
            self.params.nms_stream, self.params.nms_event, self._nms_keep = None, None, None
            return
        ev = torch.cuda.Event()
        with torch.cuda.device(self.dev):
            ev.record(stream)            # torch creates the cudaEvent lazily
        self._nms_keep = (stream, ev)
        self.params.nms_stream, self.params.nms_event = stream.cuda_stream, ev.cuda_event

    def enqueue(self, raws) -> None:
        """``raws``: the three head tensors, or (``predecoded`` plans) the one decoded ``[B, N, 5+C]`` tensor."""
        with _on(self.dev):
            self._enqueue(raws)

    def _enqueue(self, raws) -> None:
        L = _lib.lib()
        p = self.params
        if self.predecoded:
            preds = _req(raws[0] if isinstance(raws, (list, tuple)) else raws, "preds")
            if tuple(preds.shape) != (self.B, self.N, p.C + 5 + p.extra_cols):
                raise RuntimeError("post_process: preds shape differs from the plan")
            self.raws = [preds]
        else:
            self.raws = [_req(r, f"raw[{i}]") for i, r in enumerate(raws)]
            if [tuple(r.shape) for r in self.raws] != self.shapes:
                raise RuntimeError("detect: head output shapes differ from the plan")
        if any(r.device != self.dev for r in self.raws):
            raise RuntimeError(f"detect: the plan lives on {self.dev}, the inputs on {self.raws[0].device}")
        self.mask_bytes = _mask_budget.get(self.key, DEFAULT_MASK_BYTES)
        nk = (self.mask_bytes, p.nms_path)
        if self._need_key != nk:     # (the size query is a host-side carve: remembered per scratch budget / engine)
            self._need = L.bg_detect_workspace_bytes(C.byref(p), self.mask_bytes)
            self._need_key = nk
        need = self._need
        if need == 0:
            raise RuntimeError("detect: invalid parameters")
        ws = _workspace(self.dev, self.ws_tag, need)
        if self.host_result:
            self._seq = self._seq % 0x7fffffff + 1      # a fresh value per call: the word still holds the previous one
            p.host_flag_value = self._seq
        if self.predecoded:
            check(L.bg_post_process(self.raws[0].data_ptr(), C.byref(p), self.out_boxes.data_ptr(), self.out_img.data_ptr(),
                                    self.out_keep.data_ptr(), self.counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                    self.mask_bytes, _stream(self.dev)), "bg_post_process")
            return
        rc = L.bg_detect(self.raws[0].data_ptr(), self.raws[1].data_ptr(), self.raws[2].data_ptr(), C.byref(p),
                         *self._out_ptrs, ws.data_ptr(), ws.numel(), self.mask_bytes, _stream(self.dev))
        if rc:
            check(rc, "bg_detect")

    def result(self) -> Detections:
        with _on(self.dev):
            return self._result()

    def _wait_host_flag(self) -> None:
        """Spin on the word the last kernel of the call stores in page-locked host memory (no CUDA call on the way);
        every ~2 ms make sure the stream has not finished without it (a failed launch)."""
        hdr, i, seq = self._np_hdr, self._flag_idx, self._seq
        spins = 0
        while hdr[i] != seq:
            spins += 1
            if spins % 20000 == 0 and torch.cuda.current_stream(self.dev).query() and hdr[i] != seq:
                torch.cuda.current_stream(self.dev).synchronize()
                if hdr[i] != seq:
                    raise RuntimeError("detect: the kernels finished without storing the host flag")

    def _result(self) -> Detections:
        B = self.B
        while True:
            if self.host_result:
                self._wait_host_flag()
                h = self.counts
            else:
                h = _read_counts(self.counts, "detect")
            status = int(h[1])
            if not status and not _nms_path_hint:   # the common case, with as few tensor operations as possible (batch-1 latency)
                hc = h.clone()
                k = int(hc[0])
                return Detections(self.out_boxes[:k], self.out_img[:k], self.out_keep[:k], hc[2: 2 + B], hc[2 + B: 2 + 2 * B])
            if status & _lib.STATUS_NEED_GENERAL:
                # an image exceeded what the one-CTA-per-image NMS holds: run again through the general engine
                # and remember it (survivor overflow is re-evaluated from the counts, edge overflow is kept)
                most = int(h[2 + B: 2 + 2 * B].max())
                if self.params.nms_path == 5 and PER_IMAGE_NMS_CAP_LEAN < most <= PER_IMAGE_NMS_CAP:
                    nxt = 2       # more than the lean kernel holds: the 1024-thread kernel
                elif most > PER_IMAGE_NMS_CAP and most <= PER_IMAGE_NMS_CAP_LARGE and self.params.nms_path != 4:
                    nxt = 4       # the larger per-image kernel holds it
                elif most > PER_IMAGE_NMS_CAP_LARGE:
                    nxt = 1       # general engine while the images are this crowded
                else:
                    nxt = 3       # overlap-edge overflow: general engine for good
                _nms_path_hint[self.hint_key] = nxt
                self.params.nms_path = {2: 2, 4: 4}.get(nxt, 1)
                self._enqueue(self.raws)
                continue
            if _nms_path_hint.get(self.hint_key) in (1, 2, 4):   # crowded earlier; step back down when it is sparse again
                most = int(h[2 + B: 2 + 2 * B].max())
                if self.params.nms_path == 1 and most <= PER_IMAGE_NMS_CAP_LARGE // 2:
                    _nms_path_hint[self.hint_key] = 4 if most > PER_IMAGE_NMS_CAP // 2 else 0
                elif self.params.nms_path == 4 and most <= PER_IMAGE_NMS_CAP // 2:
                    _nms_path_hint.pop(self.hint_key, None)
                elif self.params.nms_path == 2 and _nms_path_hint.get(self.hint_key) == 2 and most <= PER_IMAGE_NMS_CAP_LEAN // 2:
                    _nms_path_hint.pop(self.hint_key, None)
            if int(h[1]) & _lib.STATUS_MASK_SPACE:
                # neither the overlap-edge list nor the dense bit matrix fitted.  The per-image survivor counts are
                # known now: grow the scratch fourfold (room for more edges) up to the exact matrix size, run again
                cand = h[2 + B: 2 + 2 * B].to(torch.int64)
                exact = int((cand * ((cand + 63) // 64)).sum()) * 8
                if exact <= self.mask_bytes:
                    raise RuntimeError("detect: suppression-mask workspace exhausted")
                _mask_budget[self.key] = min(exact + (exact >> 3), self.mask_bytes * 4)
                self._enqueue(self.raws)
                continue
            k = int(h[0])
            return Detections(self.out_boxes[:k], self.out_img[:k], self.out_keep[:k], h[2: 2 + B].clone(),
                              h[2 + B: 2 + 2 * B].clone())


    def decode_rows(self, idx: torch.Tensor) -> torch.Tensor:
        """``[K, 5+C(+extra)]`` rows ``[obj logit, class logits, x, y, w, h, ...]`` of the flat candidates ``idx`` (e.g.
        ``Detections.keep_idxs``) of the last enqueued batch: what ``DetectionNet.forward(x, inference=True)`` holds for
        them (modules/detection.py:69-91), decoded exactly as the fused kernels decode internally."""
        if self.predecoded:
            raise RuntimeError("decode_rows: the plan was built on already decoded rows")
        idx = _req(idx, "idx", torch.int64).reshape(-1)
        D = self.params.C + 5 + self.params.extra_cols
        out = torch.empty(idx.numel(), D, dtype=torch.float32, device=self.dev)
        with _on(self.dev):
            check(_lib.lib().bg_decode_rows(self.raws[0].data_ptr(), self.raws[1].data_ptr(), self.raws[2].data_ptr(),
                                            C.byref(self.params), idx.data_ptr(), idx.numel(), out.data_ptr(), _stream(self.dev)),
                  "bg_decode_rows")
        return out

    # ---- the step after the path (inference_det.py:100-129) ----------------------------------------
    def enqueue_host_copy(self) -> None:
        """Queue ONE device->host copy of [counts | rows] behind the kernels of the last ``enqueue`` (pinned buffer;
        sized for the row count of recent batches plus head room, the rare overflow is fetched by ``result_host``)."""
        if self.host_result:
            return                              # the kernels write the host buffer themselves
        if self._host_packed is None:
            self._host_packed = torch.empty(self._packed.numel(), dtype=torch.uint8).pin_memory()
        n = min(self._packed.numel(), self._hdr_bytes + self._rows_guess * 24)
        with _on(self.dev):
            self._host_packed[:n].copy_(self._packed[:n], non_blocking=True)
        self._copied_rows = (n - self._hdr_bytes) // 24

    def result_host(self) -> "HostDetections":
        """Rows of the last batch on the HOST, image by image, without ``.unique()``, per-image masks or per-image
        ``.cpu()`` calls: what the reference's host loop builds at inference_det.py:100-129 (``boxes.detach().cpu().numpy()``
        per image, after the tracked-class filter), from one pinned buffer filled by one asynchronous copy."""
        import numpy as np
        B = self.B
        if self.host_result:
            if self.params.order != 0:
                raise RuntimeError("result_host: needs order='image' (rows grouped by image)")
            self._wait_host_flag()
            hdr = self._np_hdr
            if hdr[1]:                      # a status bit (rare): the general route re-runs on the right engine
                with _on(self.dev):
                    self._result()
            k = int(hdr[0])
            offsets = np.zeros(B + 1, np.int64)
            np.cumsum(hdr[2: 2 + B], out=offsets[1:])
            return HostDetections(self._np_rows[:k], offsets)
        if getattr(self, "_copied_rows", None) is None:
            self.enqueue_host_copy()
        with _on(self.dev):
            torch.cuda.current_stream(self.dev).synchronize()
            hdr = self._host_packed[: (2 + 2 * B) * 4].view(torch.int32)
            if int(hdr[1]) & (_lib.STATUS_NEED_GENERAL | _lib.STATUS_MASK_SPACE):
                self._result()              # re-runs on the right engine (rare); then copy again
                self._copied_rows = None
                self.enqueue_host_copy()
                torch.cuda.current_stream(self.dev).synchronize()
                hdr = self._host_packed[: (2 + 2 * B) * 4].view(torch.int32)
            k = int(hdr[0])
            if k > self._copied_rows:       # more rows than the optimistic copy covered: fetch the rest
                lo, hi = self._hdr_bytes + self._copied_rows * 24, self._hdr_bytes + k * 24
                self._host_packed[lo:hi].copy_(self._packed[lo:hi], non_blocking=True)
                torch.cuda.current_stream(self.dev).synchronize()
        self._rows_guess = max(256, k + k // 4)
        self._copied_rows = None
        rows = self._host_packed[self._hdr_bytes: self._hdr_bytes + k * 24].view(torch.float32).view(k, 6).numpy()
        counts = hdr[2: 2 + B].numpy().astype(np.int64)
        offsets = np.zeros(B + 1, np.int64)
        np.cumsum(counts, out=offsets[1:])
        if self.params.order != 0:
            raise RuntimeError("result_host: needs order='image' (rows grouped by image)")
        return HostDetections(rows, offsets)


@dataclass
class HostDetections:
    """``rows [K, 6]`` float32 on the host = (score, class, x1, y1, x2, y2), image-major and score-descending inside an
    image; ``offsets [B+1]``: rows of image b are ``rows[offsets[b]:offsets[b+1]]`` (CSR).  The arrays are views of
    the plan's pinned buffer: consume them before the plan's next ``result_host``."""
    rows: "object"
    offsets: "object"

    def per_image(self):
        """Yields ``(image index, boxes [k, 6])`` for the images that kept at least one row -- the iteration of
        inference_det.py:100-113 (``sample_idxs.unique()`` + mask + tracked-class filter + ``.cpu().numpy()``)."""
        for b in range(len(self.offsets) - 1):
            lo, hi = int(self.offsets[b]), int(self.offsets[b + 1])
            if hi > lo:
                yield b, self.rows[lo:hi]


class DetectPipeline:
    """Consecutive batches of one fused decode+NMS configuration in flight on ``depth`` CUDA streams, each with its
    own plan, scratch and output buffers.  The decode kernel is HBM-bound and the per-image NMS is a latency-bound
    tail on a subset of the SMs, so the next batch's decode fills the machine while the previous batch resolves.

    ``submit(raws)`` enqueues a batch (after whatever the caller's current stream has queued) and returns its slot;
    ``result(slot)`` is that batch's :class:`Detections` (one host read).  A slot's buffers are reused ``depth``
    submissions later: take the result before that."""

    def __init__(self, shapes, anchors3, input_shape, num_classes, device, og_size=None, iou_threshold=0.5,
                 score_threshold=0.1, box_allowance=None, tracked_classes=None, order="image", variant=0,
                 depth: int = 4, nms_priority: bool = True):
        if depth < 1:
            raise RuntimeError("DetectPipeline: depth must be at least 1")
        self.depth = int(depth)
        self.plans = []
        self.streams = [torch.cuda.Stream(device=device) for _ in range(self.depth)]
        # the NMS kernels of a batch go to a second, higher-priority stream: the block scheduler then places the batch's few
        # NMS CTAs (which fit next to resident decode CTAs) ahead of the pending decode CTAs of the batches behind it
        self.nms_streams = [torch.cuda.Stream(device=device, priority=-1) for _ in range(self.depth)] if nms_priority and self.depth > 1 else []
        for i in range(self.depth):
            pl = DetectPlan(shapes, anchors3, input_shape, num_classes, device, og_size, iou_threshold, score_threshold,
                            box_allowance, tracked_classes, order, variant, "auto", False, throughput=self.depth > 1)
            pl.ws_tag = "detect/pipe%d" % i
            if self.nms_streams:
                pl.use_nms_stream(self.nms_streams[i])
            self.plans.append(pl)
        self.submitted = 0

    def submit(self, raws) -> int:
        slot = self.submitted % self.depth
        st = self.streams[slot]
        st.wait_stream(torch.cuda.current_stream(st.device))
        with torch.cuda.stream(st):
            self.plans[slot].enqueue(raws)
        self.submitted += 1
        return slot

    def result(self, slot: int) -> Detections:
        with torch.cuda.stream(self.streams[slot]):
            return self.plans[slot].result()

    def join(self) -> None:
        """Make the caller's current stream wait for every batch submitted so far."""
        cur = torch.cuda.current_stream(self.streams[0].device)
        for st in self.streams:
            cur.wait_stream(st)


def detect(raws: Sequence[torch.Tensor], anchors3: Sequence, input_shape: Tuple[int, int], num_classes: int,
           og_size: Optional[Tuple[int, int]] = None, iou_threshold: float = 0.5, score_threshold: float = 0.1,
           box_allowance: Optional[float] = None, tracked_classes: Optional[Sequence[int]] = None,
           order: str = "image", variant: int = 0, nms_path: str = "auto") -> Detections:
    """Fused ``DetectionNet.forward(inference=True)`` tail (modules/detection.py:69-91) +
    ``post_process_preds`` lines 57-97 and the class filter at :107-109, from the three raw head outputs
    ``[B, ny, nx, na, 5+C]``.  The returned tensors are views of buffers owned by the plan of this call."""
    raws = [_req(r, f"raw[{i}]") for i, r in enumerate(raws)]
    plan = DetectPlan([tuple(r.shape) for r in raws], anchors3, input_shape, num_classes, raws[0].device, og_size,
                      iou_threshold, score_threshold, box_allowance, tracked_classes, order, variant, nms_path)
    plan.enqueue(raws)
    return plan.result()


def extra_columns(preds: torch.Tensor, det: Detections, num_classes: int) -> torch.Tensor:
    """Trailing columns of the kept rows, ``preds[..., 5+num_classes:]`` gathered by ``det.keep_idxs`` -- the
    ``mask_coefs`` / ``keypoints`` of ``inference_seg.post_process_preds`` (lines 66-68, 94-95) in the row order of
    ``det.pred_boxes``."""
    flat = preds.reshape(-1, preds.shape[-1])
    return flat[det.keep_idxs, 5 + num_classes:]


def seg_masks(coefs: torch.Tensor, counts, protos: torch.Tensor, out_size: Tuple[int, int]) -> torch.Tensor:
    """``inference_seg.post_process_preds`` lines 115-117 for every kept row at once: ``sigmoid(coefs @ protos_i)`` on the
    protos' grid, bilinear resize (``align_corners=False``) to ``out_size``, ``> 0.5``.  ``coefs [n, K]``: the coefficient
    columns of the kept rows, image by image (``extra_columns(...)[:, :K]`` of an ``order="image"`` result); ``counts [B]``:
    rows per image (``Detections.counts``); ``protos [B, K, Hp, Wp]``.  Returns ``bool [n, H, W]`` on the device."""
    coefs = _req(coefs, "coefs")
    protos = _req(protos, "protos")
    if coefs.dim() != 2 or protos.dim() != 4 or coefs.shape[1] != protos.shape[1]:
        raise RuntimeError("seg_masks: coefs [n, K] and protos [B, K, Hp, Wp] do not fit")
    dev = _same_device(coefs, protos)
    B, K, Hp, Wp = (int(v) for v in protos.shape)
    n, (H, W) = int(coefs.shape[0]), (int(out_size[0]), int(out_size[1]))
    cnt = torch.as_tensor(counts, dtype=torch.int64).reshape(-1).cpu()
    if cnt.numel() != B or int(cnt.sum()) != n:
        raise RuntimeError("seg_masks: counts must hold the rows of each of the B images")
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(cnt, 0)
    out = torch.empty(n, H, W, dtype=torch.uint8, device=dev)
    if n == 0:
        return out.view(torch.bool)
    with _on(dev):
        off_d = off.to(dev, non_blocking=True)
        low = torch.empty(n, Hp * Wp, dtype=torch.float32, device=dev)
        check(_lib.lib().bg_seg_masks(coefs.data_ptr(), off_d.data_ptr(), protos.data_ptr(), B, K, Hp, Wp, n, H, W, low.data_ptr(),
                                      out.data_ptr(), _stream(dev)), "bg_seg_masks")
    return out.view(torch.bool)


def post_process(preds: torch.Tensor, input_shape: Tuple[int, int], num_classes: int, iou_threshold: float = 0.5,
                 score_threshold: float = 0.1, box_allowance: Optional[float] = None,
                 tracked_classes: Optional[Sequence[int]] = None, order: str = "global", na: int = 3,
                 strides: Sequence[int] = (8, 16, 32), nms_path: str = "auto") -> Detections:
    """The compute of ``inference_det.post_process_preds`` (lines 57-97 and 107-109) on the tensor the reference
    hands it: ``preds [B, N, 5+C] = DetectionNet.forward(x, inference=True)`` with rows ``[obj, cls*C, x, y, w, h]``
    (logits + decoded pixel boxes; the segmentation / keypoint heads append mask coefficients and keypoints,
    ``inference_seg.py:58,62-97`` -- the same arithmetic; fetch those columns with :func:`extra_columns`).
    Scores, box allowance, xyxy, per-image NMS, strict score threshold, rows
    ``(score, class, x1, y1, x2, y2)`` and the tracked-class filter in one pass; ``order='global'`` is the
    reference's row order (score-descending over the batch)."""
    preds = _req(preds, "preds")
    if preds.dim() != 3 or preds.shape[2] < num_classes + 5:
        raise RuntimeError("post_process: expected preds [B, N, 5 + num_classes (+ mask coefficients / keypoints)]")
    B, N, D = preds.shape
    H, W = int(input_shape[0]), int(input_shape[1])
    shapes = [(B, H // s, W // s, na, D) for s in strides]
    if sum(sh[1] * sh[2] * na for sh in shapes) != N:
        raise RuntimeError("post_process: N does not match input_shape / strides / na")
    anchors3 = [[[1.0, 1.0]] * na] * 3  # unused: the boxes are already decoded
    plan = DetectPlan(shapes, anchors3, (H, W), num_classes, preds.device, None, iou_threshold, score_threshold,
                      box_allowance, tracked_classes, order, 0, nms_path, predecoded=True)
    plan.enqueue(preds)
    return plan.result()


def decode_scale(scale_pred: torch.Tensor, anchors, input_shape: Tuple[int, int], inference: bool = False,
                 og_size: Optional[Tuple[int, int]] = None, num_classes: Optional[int] = None, tanh_cols: int = 0) -> torch.Tensor:
    """``DetectionNet._get_scale_pred`` (modules/detection.py:98-173), optionally followed by
    ``_bbox_to_size`` (:175-190, guard of :76 applied inside).  ``num_classes`` (default: row length - 5) tells where the
    box sits when the rows carry more columns behind it; the first ``tanh_cols`` of those are the segmentation head's mask
    coefficients, which the reference passes through ``tanh`` (:131-134)."""
    x = _req(scale_pred, "scale_pred")
    B, ny, nx, na, D = x.shape
    Cc = D - 5 if num_classes is None else int(num_classes)
    extra = D - 5 - Cc
    if Cc <= 0 or extra < 0 or not 0 <= tanh_cols <= extra:
        raise RuntimeError("decode_scale: rows of %d columns do not hold 5 + %d columns (+ %d through tanh)" % (D, Cc, tanh_cols))
    out = torch.empty_like(x)
    og = (int(og_size[0]), int(og_size[1])) if og_size is not None else (-1, -1)
    with _on(x.device):
        check(_lib.lib().bg_decode_scale_ex(x.data_ptr(), out.data_ptr(), B, ny, nx, na, Cc, extra, int(tanh_cols),
                                            _anchor_array(anchors), int(input_shape[0]), int(input_shape[1]),
                                            int(bool(inference)), og[0], og[1], _stream(x.device)), "bg_decode_scale")
    return out


class _TrainDecode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, num_classes, tanh_cols):
        ctx.save_for_backward(x)
        ctx.cfg = (num_classes, tanh_cols)
        return decode_scale(x, [[1.0, 1.0]] * x.shape[3], (1, 1), False, None, num_classes, tanh_cols)

    @staticmethod
    def backward(ctx, go):
        (x,) = ctx.saved_tensors
        Cc, tc = ctx.cfg
        go = go.contiguous() if go.dtype == torch.float32 else go.float().contiguous()
        gr = torch.empty_like(x)
        with _on(x.device):
            check(_lib.lib().bg_decode_train_bwd_ex(x.data_ptr(), go.data_ptr(), gr.data_ptr(), x.numel() // x.shape[-1],
                                                    Cc, x.shape[-1] - 5 - Cc, tc, _stream(x.device)), "bg_decode_train_bwd")
        return gr, None, None


def decode_train(scale_pred: torch.Tensor, num_classes: Optional[int] = None, tanh_cols: int = 0) -> torch.Tensor:
    """``DetectionNet._get_scale_pred(..., inference=False)`` (modules/detection.py:98-173), differentiable: one CUDA
    kernel each way instead of ~15 ATen kernels and their autograd graph.  (The fused loss does not need it -- it
    takes the logits, ``detection_loss(..., input_form="raw")``; this is for every other consumer.)  ``num_classes`` /
    ``tanh_cols``: rows with more columns behind the box, the first ``tanh_cols`` of them through ``tanh`` (the segmentation
    head's mask coefficients, :131-134)."""
    x = _req(scale_pred, "scale_pred")
    Cc = x.shape[-1] - 5 if num_classes is None else int(num_classes)
    return _TrainDecode.apply(x, Cc, int(tanh_cols))


def bbox_to_size(pred: torch.Tensor, _from: torch.Tensor, _to: torch.Tensor, num_classes: int) -> torch.Tensor:
    """``DetectionNet._bbox_to_size`` (modules/detection.py:175-190): rescales the box columns of decoded rows in
    place and returns ``pred.contiguous()`` like the reference.  ``_from`` / ``_to`` are the int64 device tensors the
    reference builds at :77-78 and are read on the device."""
    if not pred.is_cuda or pred.dtype != torch.float32:
        raise RuntimeError("bbox_to_size: the box-geometry kernels need a CUDA fp32 tensor (no CPU fallback exists)")
    out = pred if pred.is_contiguous() else pred.contiguous()
    D = out.shape[-1]
    f = _from.to(device=out.device, dtype=torch.int64).contiguous()
    t = _to.to(device=out.device, dtype=torch.int64).contiguous()
    if f.numel() != 4 or t.numel() != 4:
        raise RuntimeError("bbox_to_size: _from / _to must hold four values")
    with _on(out.device):
        check(_lib.lib().bg_bbox_to_size(out.data_ptr(), out.numel() // D, int(num_classes), D, f.data_ptr(), t.data_ptr(),
                                         _stream(out.device)), "bg_bbox_to_size")
    if out is not pred:
        pred.copy_(out)  # the reference writes through `pred` (a view) before returning the contiguous copy
    return out


# ---------------------------------------------------------------------------------------------- B1
def build_target_by_scale(targets: torch.Tensor, fmap_shape, anchors, anchor_threshold: float = 4.0,
                          edge_threshold: float = 0.5, overlap_masks: Optional[bool] = None,
                          batch_size: Optional[int] = None):
    """``DetectionDataset.build_target_by_scale`` (dataset/detection_dataset.py:90-246), including the
    segmentation (``overlap_masks``) and keypoint-column variants.  Returns
    ``(indices, classes, anchors, boxes, tmask_idx, keypoints)`` exactly like the reference."""
    t = _req(targets, "targets")
    if t.dim() != 2 or t.shape[1] < 6:
        raise RuntimeError("build_target_by_scale: targets must be [nt, 6 + keypoint columns]")
    if overlap_masks and not batch_size:
        raise ValueError("batch_size is required when overlap_mask is set to True")  # the reference's own error (:149-150)
    with _on(t.device):
        return _build_target_on(t, fmap_shape, anchors, anchor_threshold, edge_threshold, overlap_masks, batch_size)


def _build_target_on(t, fmap_shape, anchors, anchor_threshold, edge_threshold, overlap_masks, batch_size):
    dev = t.device
    nt, stride = t.shape
    ny, nx = (int(v) for v in (fmap_shape.tolist() if isinstance(fmap_shape, torch.Tensor) else fmap_shape))
    anc = _anchors_host(anchors)
    na = len(anc)
    cap = max(5 * na * nt, 1)
    idx4 = torch.empty(4, cap, dtype=torch.int64, device=dev)
    cls = torch.empty(cap, dtype=torch.int64, device=dev)
    anc_out = torch.empty(cap, 2, dtype=torch.float32, device=dev)
    box = torch.empty(cap, 4, dtype=torch.float32, device=dev)
    count = torch.empty(2, dtype=torch.int32, device=dev)
    L = _lib.lib()
    plain = overlap_masks is None and stride == 6
    if plain:
        ws = _workspace(dev, "assign", max(L.bg_assign_workspace_bytes(nt, na), 256))
        check(L.bg_assign_targets(t.data_ptr(), nt, ny, nx, _anchor_array(anc), na, float(anchor_threshold),
                                  float(edge_threshold), idx4.data_ptr(), cls.data_ptr(), anc_out.data_ptr(),
                                  box.data_ptr(), cap, count.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)),
              "bg_assign_targets")
        M = int(_read_counts(count[:1], "assign")[0])
        tmask = kpts = None
    else:
        mode = 0 if overlap_masks is None else (2 if overlap_masks else 1)
        bs = int(batch_size) if batch_size else 0
        tmask = torch.empty(cap, dtype=torch.int64, device=dev) if mode else None
        kpts = torch.empty(cap, stride - 6, dtype=torch.float32, device=dev) if stride > 6 else None
        ws = _workspace(dev, "assign", max(L.bg_assign_ex_workspace_bytes(nt, na, bs), 256))
        check(L.bg_assign_targets_ex(t.data_ptr(), nt, stride, ny, nx, _anchor_array(anc), na, float(anchor_threshold),
                                     float(edge_threshold), mode, bs, idx4.data_ptr(), cls.data_ptr(), anc_out.data_ptr(),
                                     box.data_ptr(), tmask.data_ptr() if mode else None,
                                     kpts.data_ptr() if kpts is not None else None, cap, count.data_ptr(), ws.data_ptr(),
                                     ws.numel(), _stream(dev)), "bg_assign_targets_ex")
        h = _read_counts(count, "assign")
        if int(h[1]):
            raise RuntimeError("build_target_by_scale: per-image target counts do not add up to the number of targets "
                               "(image ids outside 0..batch_size-1)")
        M = int(h[0])
        tmask = tmask[:M] if tmask is not None else None
        kpts = kpts[:M] if kpts is not None else None
    indices = [idx4[k, :M] for k in range(4)]
    return indices, cls[:M], anc_out[:M], box[:M], tmask, kpts


# ---------------------------------------------------------------------------------------------- B2
class _CIoU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t, e):
        p32, t32 = _req(p, "preds_xywh").reshape(-1, 4), _req(t, "targets_xywh").reshape(-1, 4)
        dev = _same_device(p32, t32)
        out = torch.empty(p32.shape[0], dtype=torch.float32, device=dev)
        with _on(dev):
            check(_lib.lib().bg_ciou_fwd(p32.data_ptr(), t32.data_ptr(), p32.shape[0], float(e), out.data_ptr(),
                                         _stream(dev)), "bg_ciou_fwd")
        ctx.save_for_backward(p32, t32)
        ctx.e = float(e)
        ctx.shape = p.shape
        return out.reshape(p.shape[:-1])

    @staticmethod
    def backward(ctx, go):
        p32, t32 = ctx.saved_tensors
        go = go.contiguous().reshape(-1).float()
        gp = torch.empty_like(p32)
        with _on(p32.device):
            check(_lib.lib().bg_ciou_bwd(p32.data_ptr(), t32.data_ptr(), go.data_ptr(), p32.shape[0], ctx.e,
                                         gp.data_ptr(), _stream(p32.device)), "bg_ciou_bwd")
        return gp.reshape(ctx.shape), None, None


def compute_ciou(preds_xywh: torch.Tensor, targets_xywh: torch.Tensor, e: float = 1e-7) -> torch.Tensor:
    """``DetectionLoss.compute_ciou`` (modules/detection_loss.py:229-264), differentiable w.r.t. ``preds_xywh`` (alpha held
    constant as under the reference's ``no_grad``).  Element-wise form, or the broadcasting form of :231-234
    (``preds [..., A, 4]`` against ``targets [..., 4]``: every target against the ``A`` boxes of its row)."""
    if preds_xywh.dim() == targets_xywh.dim() + 1:
        targets_xywh = targets_xywh.unsqueeze(-2)
    if preds_xywh.dim() != targets_xywh.dim() or preds_xywh.shape[-1] != 4 or targets_xywh.shape[-1] != 4:
        raise RuntimeError("compute_ciou: expected preds [..., 4] and targets of the same rank or one dimension less")
    if preds_xywh.shape != targets_xywh.shape:
        if targets_xywh.requires_grad:
            raise RuntimeError("compute_ciou: gradients flow to preds_xywh only; detach the broadcast targets")
        shape = torch.broadcast_shapes(preds_xywh.shape, targets_xywh.shape)
        if tuple(shape) != tuple(preds_xywh.shape):
            preds_xywh = preds_xywh.expand(shape)
        targets_xywh = targets_xywh.expand(shape)
    return _CIoU.apply(preds_xywh, targets_xywh, e)


# ---------------------------------------------------------------------------------------------- B3
METRIC_KEYS = ("mean_ciou", "conf_loss", "avg_pos_conf", "avg_neg_conf", "class_loss", "accuracy", "f1", "precision",
               "recall")


_loss_param_cache: Dict[tuple, LossParams] = {}
_FORMS = {"decoded": _lib.LOSS_DECODED, "raw": _lib.LOSS_RAW, "split": _lib.LOSS_RAW_SPLIT}


def _loss_params(shapes, C_cls, extra, nt, anchors3, cfg, form) -> LossParams:
    """The parameter block of bg_loss_fwd / bg_loss_bwd; built once per (shapes, target count, anchors, weights)."""
    akey = tuple(tuple(map(tuple, _anchors_host(a))) for a in anchors3)   # by value
    sw = cfg.get("scale_w") or [4.0, 2.0, 1.0]
    key = (shapes, C_cls, extra, nt, akey, cfg.get("anchor_t", 4.0), cfg.get("edge_t", 0.5), cfg.get("label_smoothing", 0.0),
           cfg.get("box_w", 1.0), cfg.get("conf_w", 1.0), cfg.get("class_w", 1.0), tuple(sw), form)
    hit = _loss_param_cache.get(key)
    if hit is not None:
        return hit
    if len(_loss_param_cache) > 256:
        _loss_param_cache.clear()
    p = _loss_param_cache[key] = LossParams()
    B, _, _, na = shapes[0]
    p.B, p.C, p.na = B, C_cls, na
    for s, sh in enumerate(shapes):
        if sh[0] != B or sh[3] != na:
            raise RuntimeError("detection_loss: inconsistent prediction shapes")
        p.ny[s], p.nx[s] = sh[1], sh[2]
        anc = _anchors_host(anchors3[s])
        if len(anc) != na:
            raise RuntimeError("detection_loss: anchors do not match the prediction tensors")
        for a, (w, h) in enumerate(anc):
            p.anchors[s][a][0] = w
            p.anchors[s][a][1] = h
    p.anchor_t, p.edge_t = float(cfg.get("anchor_t", 4.0)), float(cfg.get("edge_t", 0.5))
    p.label_smoothing = float(cfg.get("label_smoothing", 0.0))
    p.box_w, p.conf_w, p.class_w = float(cfg.get("box_w", 1.0)), float(cfg.get("conf_w", 1.0)), float(cfg.get("class_w", 1.0))
    for s in range(3):
        p.scale_w[s] = float(sw[s])
    p.nt = nt
    p.input_form = form
    p.extra_cols = extra
    return p


def _head_ptrs(tensors, split: bool):
    arr = _lib.HeadPtrs3()
    if split:
        for s in range(3):
            arr[s].obj, arr[s].cls, arr[s].box = (tensors[3 * s].data_ptr(), tensors[3 * s + 1].data_ptr(),
                                                  tensors[3 * s + 2].data_ptr())
    else:
        for s in range(3):
            arr[s].obj = tensors[s].data_ptr()
    return arr


_side_streams: Dict[int, "torch.cuda.Stream"] = {}


def _side_stream(dev: torch.device, high_priority: bool = False) -> "torch.cuda.Stream":
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _side_streams.get((idx, high_priority))
    if st is None:
        st = _side_streams[(idx, high_priority)] = torch.cuda.Stream(device=dev, priority=-1 if high_priority else 0)
    return st


def _split_grad_buffers(tensors, dev):
    """Gradient tensors of the split form: the class / box planes of the three scales share one buffer (one memset
    clears them all), the objectness planes are separate."""
    big = [t for i, t in enumerate(tensors) if i % 3]
    flat = torch.empty(sum(t.numel() for t in big), dtype=torch.float32, device=dev)
    grads, off = [], 0
    for i, t in enumerate(tensors):
        if i % 3 == 0:
            grads.append(torch.empty_like(t))
        else:
            grads.append(flat[off: off + t.numel()].view(t.shape))
            off += t.numel()
    return grads, flat


class _DetLoss(torch.autograd.Function):
    """Forward: assignment + gather + CIoU + objectness/class BCE for the three scales and the combined loss,
    all on the device, no host sync.  Backward: dense gradients written once per scale; the upstream
    gradient stays on the device.  Everything the backward needs lives in a workspace allocated per forward and
    owned by ``ctx`` -- any number of forwards may precede their backwards (gradient accumulation, several
    loss modules, a validation loss in between).

    Split form, optional (``PRECLEAR_SPLIT_GRADS``): the 2 GB of zeros of the class / box gradient planes do not depend
    on the forward; the gradient tensors can be allocated already in the forward and cleared on a second stream next
    to the forward kernels, the backward then joins that stream and only writes the objectness plane and the matched
    rows."""

    @staticmethod
    def forward(ctx, targets, params: LossParams, scalars, hist, status, *tensors):
        L = _lib.lib()
        dev = tensors[0].device
        split = params.input_form == _lib.LOSS_RAW_SPLIT
        ctx.pre = None
        with _on(dev):
            preclear = split and PRECLEAR_SPLIT_GRADS and any(ctx.needs_input_grad[5:])   # (whether a backward can follow)
            fwd_stream = _stream(dev)
            hp = None
            if preclear and PRECLEAR_SPLIT_GRADS == "priority":
                # the clear on the caller's stream, the forward kernels on a HIGHER-PRIORITY stream next to it: the block
                # scheduler serves their CTAs ahead of the memset's pending ones, so the latency-bound forward runs inside
                # the bandwidth-bound clear instead of queueing behind it
                cur, hp = torch.cuda.current_stream(dev), _side_stream(dev, True)
                grads, flat = _split_grad_buffers(tensors, dev)
                ws = torch.empty(L.bg_loss_workspace_bytes(C.byref(params)), dtype=torch.uint8, device=dev)
                loss = torch.empty(1, dtype=torch.float32, device=dev)
                hp.wait_stream(cur)
                check(L.bg_loss_clear_grads(C.byref(params), _head_ptrs(grads, True), cur.cuda_stream), "bg_loss_clear_grads")
                ctx.pre = (grads, flat, None)
                fwd_stream = hp.cuda_stream
            elif preclear:
                cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
                grads, flat = _split_grad_buffers(tensors, dev)
                side.wait_stream(cur)        # the buffers may be recycled memory of work queued on this stream
                check(L.bg_loss_clear_grads(C.byref(params), _head_ptrs(grads, True), side.cuda_stream), "bg_loss_clear_grads")
                if not torch.cuda.is_current_stream_capturing():
                    flat.record_stream(side)   # (if the node dies without its backward, the memory waits for the clear)
                ctx.pre = (grads, flat, side)
            if hp is None:
                ws = torch.empty(L.bg_loss_workspace_bytes(C.byref(params)), dtype=torch.uint8, device=dev)
                loss = torch.empty(1, dtype=torch.float32, device=dev)
            check(L.bg_loss_fwd(_head_ptrs(tensors, split), targets.data_ptr() if targets.numel() else None,
                                C.byref(params), scalars.data_ptr(), hist.data_ptr(), loss.data_ptr(), status.data_ptr(),
                                ws.data_ptr(), ws.numel(), fwd_stream), "bg_loss_fwd")
            if hp is not None:
                torch.cuda.current_stream(dev).wait_stream(hp)   # loss, scalars and the cleared planes: all behind this point
        ctx.save_for_backward(*tensors)
        ctx.params, ctx.ws = params, ws
        return loss.reshape(())

    @staticmethod
    def backward(ctx, go):
        tensors = ctx.saved_tensors
        L = _lib.lib()
        params = ctx.params
        dev = tensors[0].device
        split = params.input_form == _lib.LOSS_RAW_SPLIT
        flags = 0
        with _on(dev):
            if ctx.pre is not None:
                grads, _flat, side = ctx.pre
                if side is not None:
                    torch.cuda.current_stream(dev).wait_stream(side)   # the planes are zero from here on
                flags = _lib.LOSS_BWD_PRECLEARED
                ctx.pre = None                                     # (a second backward through the same node is refused by autograd anyway)
            elif split:
                grads, _flat = _split_grad_buffers(tensors, dev)
            else:
                grads = [torch.empty_like(x) for x in tensors]
            if go.dtype != torch.float32 or not go.is_contiguous():
                go = go.to(torch.float32).contiguous()
            check(L.bg_loss_bwd(_head_ptrs(tensors, split), C.byref(params), go.data_ptr(), 1.0, _head_ptrs(grads, split), flags,
                                ctx.ws.data_ptr(), ctx.ws.numel(), _stream(dev)), "bg_loss_bwd")
        return (None, None, None, None, None, *grads)


# split form: where the 2 GB clear of the class / box gradient planes runs.  False: inside the backward.  True: on a second
# stream next to the forward (measured on B200: gains nothing -- a memset fills every thread slot of the machine and the
# forward kernels queue behind it).  "priority": the clear on the caller's stream and the FORWARD on a higher-priority
# stream next to it, so that the scheduler serves the forward's CTAs first: graph replay 0.0916 vs 0.0997 ms at 32 images,
# 0.154 vs 0.161 at 64, 0.502 vs 0.511 at 256 (where the step sits on the DRAM floor of its 3.16 GB of actual traffic).
PRECLEAR_SPLIT_GRADS = "priority"


_combine_param_cache: Dict[tuple, LossParams] = {}


def loss_terms_pack(scalars: torch.Tensor, cells: Sequence[int], num_classes: int) -> torch.Tensor:
    """``[3, 5]`` float64 per scale ``{lbox*M, lconf*cells, lcls*M*C, M, cells}`` from the ``[3, 8]`` scalar block of
    :func:`detection_loss` on this rank's image shard: the sums that add up over shards (one tiny kernel)."""
    sc = _req(scalars, "scalars", torch.float64)
    pack = torch.empty(3, 5, dtype=torch.float64, device=sc.device)
    c3 = (C.c_int64 * 3)(*(int(c) for c in cells))
    with _on(sc.device):
        check(_lib.lib().bg_loss_pack(sc.data_ptr(), c3, int(num_classes), pack.data_ptr(), _stream(sc.device)), "bg_loss_pack")
    return pack


def loss_terms_combine(pack: torch.Tensor, cfg: dict, num_classes: int) -> torch.Tensor:
    """The loss of the concatenated batch (modules/detection_loss.py:107-110 with global means) from the summed
    terms; 0-d float64 tensor on the device."""
    pk = _req(pack, "pack", torch.float64)
    sw = tuple(cfg.get("scale_w") or [4.0, 2.0, 1.0])
    key = (int(num_classes), cfg.get("box_w", 1.0), cfg.get("conf_w", 1.0), cfg.get("class_w", 1.0), sw)
    p = _combine_param_cache.get(key)
    if p is None:
        p = _combine_param_cache[key] = LossParams()
        p.C = int(num_classes)
        p.box_w, p.conf_w, p.class_w = float(key[1]), float(key[2]), float(key[3])
        for s in range(3):
            p.scale_w[s] = float(sw[s])
    out = torch.empty(1, dtype=torch.float64, device=pk.device)
    with _on(pk.device):
        check(_lib.lib().bg_loss_combine(pk.data_ptr(), C.byref(p), out.data_ptr(), _stream(pk.device)), "bg_loss_combine")
    return out.reshape(())


def _macro_metrics(hist: torch.Tensor, M: int) -> Dict[str, float]:
    """sklearn accuracy / macro f1, precision, recall (modules/detection_loss.py:198-206) from the per-class
    (tp, n_true, n_pred) counters -- finished on the host in float64."""
    nan = float("nan")
    if M == 0:
        return dict(accuracy=nan, f1=nan, precision=nan, recall=nan)
    tp, nt, npred = (hist[i].double() for i in range(3))
    lab = (nt + npred) > 0
    prec = torch.where(npred > 0, tp / npred.clamp(min=1), torch.zeros_like(tp))[lab]
    rec = torch.where(nt > 0, tp / nt.clamp(min=1), torch.zeros_like(tp))[lab]
    f1 = (2 * tp / (nt + npred).clamp(min=1))[lab]
    return dict(accuracy=float(tp.sum() / M), f1=float(f1.mean()), precision=float(prec.mean()), recall=float(rec.mean()))


def detection_loss(preds3: Sequence, targets: torch.Tensor, anchors3: Sequence, cfg: dict,
                   with_metrics: bool = True, return_scalars: bool = False, input_form: str = "decoded",
                   num_classes: Optional[int] = None):
    """``DetectionLoss.forward`` (modules/detection_loss.py:84-122) for the default configuration.
    Returns ``(loss, metrics_dict)``; ``loss`` is a 0-d tensor attached to autograd through ``preds3``.

    ``input_form``:
      ``"decoded"``  ``preds3`` = what ``DetectionNet.forward(x)`` returns in training mode (the reference's contract);
      ``"raw"``      ``preds3`` = the three head outputs ``[B,ny,nx,na,5+C]`` themselves: the training-mode decode of
                     ``_get_scale_pred`` (modules/detection.py:122,125) is fused into the loss, the gradient comes back
                     with respect to the logits;
      ``"split"``    ``preds3`` = three ``(conf [B,ny,nx,na], cls [B,ny,nx,na,C], bbox [B,ny,nx,na,4])`` triples, the
                     head's conv outputs before ``EffiDecHead.forward`` concatenates them (modules/common.py:908-919).
    ``num_classes`` is only needed when interleaved rows carry trailing columns (mask coefficients / keypoints),
    which the loss skips.  ``return_scalars=True`` appends the device tensor ``[3, 8]`` float64 of per-scale terms
    (lbox, lconf, lcls, mean_ciou, avg_pos_conf, avg_neg_conf, M, n_neg) -- what ``shard.allreduce_loss_terms``
    combines across ranks.  A target row that names an image outside the batch or a class outside ``0..C-1`` makes
    the reference raise IndexError; here it is dropped on the device and the IndexError is raised when the metrics
    are read (``with_metrics=True``)."""
    form = _FORMS[input_form]
    if form == _lib.LOSS_RAW_SPLIT:
        tensors = []
        for i, tri in enumerate(preds3):
            if len(tri) != 3:
                raise RuntimeError("detection_loss: the split form takes (conf, cls, bbox) per scale")
            conf, cls, box = (_req(x, f"preds[{i}]") for x in tri)
            if conf.dim() == 5 and conf.shape[-1] == 1:
                conf = conf.reshape(conf.shape[:-1])
            if conf.dim() != 4 or cls.dim() != 5 or box.dim() != 5 or cls.shape[:4] != conf.shape or box.shape[:4] != conf.shape \
                    or box.shape[4] != 4:
                raise RuntimeError("detection_loss: expected conf [B,ny,nx,na], cls [B,ny,nx,na,C], bbox [B,ny,nx,na,4]")
            tensors += [conf, cls, box]
        shapes = tuple(tuple(tensors[3 * s].shape) for s in range(3))
        Cc, extra = int(tensors[1].shape[4]), 0
    else:
        tensors = [_req(x, f"preds[{i}]") for i, x in enumerate(preds3)]
        if len(tensors) != 3 or any(x.dim() != 5 for x in tensors):
            raise RuntimeError("detection_loss: expected three [B, ny, nx, na, 5+C] tensors")
        D = int(tensors[0].shape[4])
        Cc = int(num_classes) if num_classes is not None else D - 5
        extra = D - 5 - Cc
        if extra < 0 or any(int(x.shape[4]) != D for x in tensors):
            raise RuntimeError("detection_loss: rows must hold 5 + num_classes (+ extra) columns")
        shapes = tuple(tuple(x.shape[:4]) for x in tensors)
    targets = _req(targets, "targets")
    if targets.dim() != 2 or targets.shape[1] != 6:
        raise RuntimeError("detection_loss: keypoint targets are out of scope for the CUDA path")
    dev = _same_device(*tensors, targets)
    params = _loss_params(shapes, Cc, extra, int(targets.shape[0]), anchors3, cfg, form)
    scalars = torch.empty(3, 8, dtype=torch.float64, device=dev)
    hist = torch.empty(3, 3, Cc, dtype=torch.int64, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    loss = _DetLoss.apply(targets, params, scalars, hist, status, *tensors)
    if cfg.get("batch_scale_loss"):
        loss = loss * shapes[-1][0]
    if not with_metrics:
        return (loss, {}, scalars) if return_scalars else (loss, {})
    # one D2H copy for everything the reference fetches with ~28 .item() calls
    host = torch.cat([scalars.reshape(-1), hist.reshape(-1).double(), loss.detach().double().reshape(1),
                      status.double()]).cpu()
    if int(host[-1]):
        raise IndexError("detection_loss: a target row names an image outside the batch or a class outside "
                         "0..num_classes-1 (index out of range)")
    sc = host[:24].reshape(3, 8)
    hh = host[24:24 + 9 * Cc].reshape(3, 3, Cc).long()
    rows = []
    for s in range(3):
        M = int(sc[s, 6])
        m = dict(mean_ciou=float(sc[s, 3]), conf_loss=float(sc[s, 1]), avg_pos_conf=float(sc[s, 4]),
                 avg_neg_conf=float(sc[s, 5]), class_loss=float(sc[s, 2]) if M else float("nan"))
        m.update(_macro_metrics(hh[s], M))
        rows.append(m)
    metrics = {"aggregate_loss": float(host[-2])}
    for k in METRIC_KEYS:
        vals = [r[k] for r in rows if r[k] == r[k]]  # pandas column mean skips NaN (:117-121)
        metrics[k] = sum(vals) / len(vals) if vals else float("nan")
    return (loss, metrics, scalars) if return_scalars else (loss, metrics)


# ---------------------------------------------------------------------------------------------- f2: SegmentationLoss
SEG_METRIC_KEYS = ("mean_ciou", "conf_loss", "seg_loss", "dice_score", "avg_pos_conf", "avg_neg_conf", "class_loss", "accuracy",
                   "f1", "precision", "recall")


def _seg_params(shapes, C_cls, K, extra, nt, anchors3, cfg, protos_shape, masks_shape) -> "_lib.SegParams":
    p = _lib.SegParams()
    B, _, _, na = shapes[0]
    p.B, p.C, p.na, p.K, p.extra_cols = B, C_cls, na, K, extra
    for s, sh in enumerate(shapes):
        p.ny[s], p.nx[s] = sh[1], sh[2]
        for a, (w, h) in enumerate(_anchors_host(anchors3[s])):
            p.anchors[s][a][0] = w
            p.anchors[s][a][1] = h
    p.anchor_t, p.edge_t = float(cfg.get("anchor_t", 4.0)), float(cfg.get("edge_t", 0.5))
    p.Hp, p.Wp = int(protos_shape[2]), int(protos_shape[3])
    p.Hm, p.Wm = int(masks_shape[1]), int(masks_shape[2])
    sw = cfg.get("scale_w") or [4.0, 2.0, 1.0]
    for s in range(3):
        p.scale_w[s] = float(sw[s])
    p.seg_w = float(cfg.get("seg_w", 1.0))
    p.nt = nt
    return p


def _ptr3(tensors) -> "C.Array":
    return _lib.Ptr3(*(t.data_ptr() for t in tensors))


class _SegLoss(torch.autograd.Function):
    """``SegmentationLoss.forward`` (modules/segmentation_loss.py:26-75) on the device: the fused detection loss on the
    prediction tensors (mask coefficients ride along as extra columns) followed by the mask term, which adds
    ``seg_w * sum_s scale_w[s] * seg_loss_s`` to the loss.  Backward: the detection backward writes the dense gradients
    (zeros in the coefficient columns), the mask backward adds the coefficient gradients of the matched rows and writes
    the gradient of the protos.  Both workspaces are allocated per forward and owned by ``ctx``."""

    @staticmethod
    def forward(ctx, targets, params, sparams, scalars, hist, status, seg_scalars, seg_status, protos, masks, *tensors):
        L = _lib.lib()
        dev = tensors[0].device
        with _on(dev):
            ws = torch.empty(L.bg_loss_workspace_bytes(C.byref(params)), dtype=torch.uint8, device=dev)
            wss = torch.empty(L.bg_seg_loss_workspace_bytes(C.byref(sparams)), dtype=torch.uint8, device=dev)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            tp = targets.data_ptr() if targets.numel() else None
            check(L.bg_loss_fwd(_head_ptrs(tensors, False), tp, C.byref(params), scalars.data_ptr(), hist.data_ptr(),
                                loss.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)), "bg_loss_fwd")
            check(L.bg_seg_loss_fwd(_ptr3(tensors), tp, protos.data_ptr(), masks.data_ptr(), C.byref(sparams), loss.data_ptr(),
                                    seg_scalars.data_ptr(), seg_status.data_ptr(), wss.data_ptr(), wss.numel(), _stream(dev)),
                  "bg_seg_loss_fwd")
        ctx.save_for_backward(protos, masks, *tensors)
        ctx.params, ctx.sparams, ctx.ws, ctx.wss = params, sparams, ws, wss
        return loss.reshape(())

    @staticmethod
    def backward(ctx, go):
        protos, masks, *tensors = ctx.saved_tensors
        L = _lib.lib()
        dev = tensors[0].device
        with _on(dev):
            grads = [torch.empty_like(x) for x in tensors]
            gprotos = torch.empty_like(protos)
            if go.dtype != torch.float32 or not go.is_contiguous():
                go = go.to(torch.float32).contiguous()
            check(L.bg_loss_bwd(_head_ptrs(tensors, False), C.byref(ctx.params), go.data_ptr(), 1.0, _head_ptrs(grads, False), 0,
                                ctx.ws.data_ptr(), ctx.ws.numel(), _stream(dev)), "bg_loss_bwd")
            check(L.bg_seg_loss_bwd(_ptr3(tensors), protos.data_ptr(), masks.data_ptr(), C.byref(ctx.sparams), go.data_ptr(),
                                    _ptr3(grads), gprotos.data_ptr(), ctx.wss.data_ptr(), ctx.wss.numel(), _stream(dev)),
                  "bg_seg_loss_bwd")
        return (None, None, None, None, None, None, None, None, gprotos, None, *grads)


def segmentation_loss(preds3: Sequence[torch.Tensor], targets: torch.Tensor, protos: torch.Tensor, target_masks: torch.Tensor,
                      anchors3: Sequence, cfg: dict, num_classes: int, num_masks: int, with_metrics: bool = True):
    """``SegmentationLoss.forward`` (modules/segmentation_loss.py:26-75) for ``overlap_masks=True``, BCE losses and no
    keypoints.  ``preds3``: the training-mode tensors ``[B, ny, nx, na, 5 + C + K (+ more)]`` of ``SegmentationNet``;
    ``protos [B, K, Hp, Wp]``; ``target_masks [B, Hm, Wm]`` (pixel = 1 + position of the covering object inside its image,
    any real dtype).  Returns ``(loss, metrics_dict)`` with the reference's twelve metric keys; ``loss`` is attached to
    autograd through ``preds3`` and ``protos``."""
    tensors = [_req(x, f"preds[{i}]") for i, x in enumerate(preds3)]
    if len(tensors) != 3 or any(x.dim() != 5 for x in tensors):
        raise RuntimeError("segmentation_loss: expected three [B, ny, nx, na, 5+C+K] tensors")
    Cc, K = int(num_classes), int(num_masks)
    D = int(tensors[0].shape[4])
    extra = D - 5 - Cc
    if extra < K or any(int(x.shape[4]) != D for x in tensors):
        raise RuntimeError("segmentation_loss: rows must hold 5 + num_classes + num_masks columns")
    if K not in (8, 16, 32):
        raise RuntimeError("segmentation_loss: the CUDA path handles 8, 16 or 32 mask coefficients")
    shapes = tuple(tuple(x.shape[:4]) for x in tensors)
    B = shapes[0][0]
    targets = _req(targets, "targets")
    if targets.dim() != 2 or targets.shape[1] != 6:
        raise RuntimeError("segmentation_loss: keypoint targets are out of scope for the CUDA path")
    protos = _req(protos, "protos")
    if protos.dim() != 4 or protos.shape[0] != B or protos.shape[1] != K:
        raise RuntimeError("segmentation_loss: protos must be [B, num_masks, Hp, Wp]")
    masks = target_masks
    if not (masks.is_cuda and masks.dim() == 3 and masks.shape[0] == B):
        raise RuntimeError("segmentation_loss: target_masks must be a CUDA tensor [B, Hm, Wm] (overlap_masks=True)")
    if masks.dtype != torch.float32 or not masks.is_contiguous():
        masks = masks.to(torch.float32).contiguous()
    dev = _same_device(*tensors, targets, protos, masks)
    nt = int(targets.shape[0])
    params = _loss_params(shapes, Cc, extra, nt, anchors3, cfg, _lib.LOSS_DECODED)
    sparams = _seg_params(shapes, Cc, K, extra, nt, anchors3, cfg, protos.shape, masks.shape)
    scalars = torch.empty(3, 8, dtype=torch.float64, device=dev)
    hist = torch.empty(3, 3, Cc, dtype=torch.int64, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    seg_scalars = torch.empty(3, 2, dtype=torch.float64, device=dev)
    seg_status = torch.empty(1, dtype=torch.int32, device=dev)
    loss = _SegLoss.apply(targets, params, sparams, scalars, hist, status, seg_scalars, seg_status, protos, masks, *tensors)
    if not with_metrics:
        return loss, {}
    host = torch.cat([scalars.reshape(-1), hist.reshape(-1).double(), seg_scalars.reshape(-1), loss.detach().double().reshape(1),
                      status.double(), seg_status.double()]).cpu()
    if int(host[-1]):
        raise RuntimeError("segmentation_loss: per-image target counts do not add up to the number of targets "
                           "(image ids outside 0..batch_size-1)")
    if int(host[-2]):
        raise IndexError("segmentation_loss: a target row names an image outside the batch or a class outside "
                         "0..num_classes-1 (index out of range)")
    sc = host[:24].reshape(3, 8)
    hh = host[24:24 + 9 * Cc].reshape(3, 3, Cc).long()
    ss = host[24 + 9 * Cc: 30 + 9 * Cc].reshape(3, 2)
    rows = []
    for s in range(3):
        M = int(sc[s, 6])
        m = dict(mean_ciou=float(sc[s, 3]), conf_loss=float(sc[s, 1]), seg_loss=float(ss[s, 0]), dice_score=float(ss[s, 1]),
                 avg_pos_conf=float(sc[s, 4]), avg_neg_conf=float(sc[s, 5]), class_loss=float(sc[s, 2]) if M else float("nan"))
        m.update(_macro_metrics(hh[s], M))
        rows.append(m)
    metrics = {"aggregate_loss": float(host[-3])}
    for k in SEG_METRIC_KEYS:
        vals = [r[k] for r in rows if r[k] == r[k]]
        metrics[k] = sum(vals) / len(vals) if vals else float("nan")
    return loss, metrics


class LossStepGraph:
    """The training-loss step -- :func:`detection_loss` forward + backward, and for image-sharded runs the per-shard
    terms, their all-reduce and the big-batch loss (``shard.allreduce_loss_terms``) -- captured once in a CUDA graph
    and replayed with a single launch.  Small shards are enqueue-bound when run eagerly (seven launches, the autograd
    engine and ~0.25 ms of Python per step against ~0.13 ms of kernels at 32 images); the replay is not.

    The graph is bound to the tensors it was captured with: ``inputs`` (leaf tensors, or ``(conf, cls, bbox)`` triples
    for ``input_form="split"``) and ``targets`` keep their addresses, new values are written INTO them
    (``copy_`` / in-place producers, or a model forward captured in the same pool); the gradients appear in the
    leaves' ``.grad`` (static buffers owned by the graph), the loss in ``self.loss`` and, when ``cells`` is given, the
    loss of the concatenated batch in ``self.combined``."""

    def __init__(self, inputs: Sequence, targets: torch.Tensor, anchors3: Sequence, cfg: dict, input_form: str = "decoded",
                 cells: Optional[Sequence[int]] = None, group=None, warmup: int = 3):
        from . import shard
        self.inputs, self.targets = list(inputs), targets
        self.leaves = [q for p in self.inputs for q in (p if isinstance(p, (tuple, list)) else (p,))]
        dev = self.leaves[0].device
        self.combined = None

        aux = torch.cuda.Stream(device=dev) if cells is not None else None

        def step():
            for p in self.leaves:
                p.grad = None
            loss, _, sc = detection_loss(self.inputs, self.targets, anchors3, cfg, with_metrics=False, return_scalars=True,
                                         input_form=input_form)
            comb = None
            if cells is not None:
                # the terms are final after the forward: pack / all-reduce / combine run on a second stream next to the
                # backward (a fork and a join in the captured graph), so the collective's latency is hidden
                cur = torch.cuda.current_stream(dev)
                aux.wait_stream(cur)
                with torch.cuda.stream(aux):
                    comb = shard.allreduce_loss_terms(sc, cells, cfg, group)
                loss.backward()
                cur.wait_stream(aux)
            else:
                loss.backward()
            return loss, comb

        with _on(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            for p in self.leaves:
                p.grad = None
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.combined = step()
        self.grads = [p.grad for p in self.leaves]

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss


# ---------------------------------------------------------------------------------------------- a13
def ratio_metrics_w_extras(anchors, wh_data: torch.Tensor, threshold: float = 4.0) -> Tuple[float, float, float]:
    """``utils/make_anchors.py:27-39``: (score, best-possible-recall, anchors-above-threshold)."""
    wh = _req(wh_data, "wh_data").reshape(-1, 2)
    anc = _anchors_host(anchors)
    out = torch.empty(3, dtype=torch.float64, device=wh.device)
    with _on(wh.device):
        check(_lib.lib().bg_ratio_metrics(wh.data_ptr(), wh.shape[0], _anchor_array(anc), len(anc), float(threshold),
                                          out.data_ptr(), _stream(wh.device)), "bg_ratio_metrics")
    s, m, n = out.cpu().tolist()
    nan = float("nan")
    return (s / n if n else nan, m / n if n else nan, m)


def ratio_metrics(anchors, wh_data: torch.Tensor, threshold: float = 4.0) -> float:
    """``utils/make_anchors.py:14-25``."""
    return ratio_metrics_w_extras(anchors, wh_data, threshold)[0]
