"""Deferred head concatenation and training-mode decode: how the fused loss reaches the head's logits with zero
edits to the host code.

In the reference, ``EffiDecHead.forward`` (modules/common.py:908-919) concatenates its three conv outputs into rows
``[obj, cls*C, tx,ty,tw,th]``, ``DetectionNet.forward`` (modules/detection.py:58-96) passes each head output through
``_get_scale_pred(inference=False)`` -- ``xy = 2*sigmoid - 0.5``, ``wh = (2*sigmoid)**2`` (:122,125) over ~15 ATen
kernels that copy the whole ``[B,ny,nx,na,5+C]`` tensor -- and the trainer hands the three results straight to
``DetectionLoss.forward`` (pipeline/detection_trainer.py:178-180).  The CUDA loss applies that decode in registers
(``BG_LOSS_RAW``) and can read the three conv outputs where they are (``BG_LOSS_RAW_SPLIT``), so neither the
concatenated nor the decoded tensors need to exist.

The patched ``EffiDecHead.forward`` / ``_get_scale_pred`` therefore return a :class:`LazyRows`: a ``torch.Tensor``
subclass that *stands for* the tensor the reference would have built (same shape, dtype, device, ``requires_grad``)
but only remembers the parts it is made of.  The patched ``DetectionLoss.forward`` recognises it and feeds the parts
to the fused loss.  Any *other* use -- indexing, arithmetic, ``torch.cat``, printing, a different loss -- goes through
``__torch_function__``, which materialises the real tensor first (once: ``torch.cat`` of the parts, then the
differentiable CUDA decode ``ops.decode_train``) and then runs the requested function on it: the values every other
consumer sees are exactly the reference's.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

_T = torch.Tensor
# metadata that the stand-in shares with the tensor it stands for: answered without materialising
_META = {
    _T.shape.__get__, _T.device.__get__, _T.dtype.__get__, _T.ndim.__get__, _T.is_cuda.__get__, _T.layout.__get__,
    _T.requires_grad.__get__, _T.size, _T.dim, _T.ndimension, _T.numel, _T.nelement, _T.stride, _T.is_contiguous,
    _T.__len__, _T.is_floating_point, _T.is_complex, _T.get_device, _T.element_size,
}


class LazyRows(torch.Tensor):
    """Stands for ``torch.cat(parts, -1)`` (``decode=False``: a head output) or for
    ``DetectionNet._get_scale_pred(torch.cat(parts, -1), ..., inference=False)`` (``decode=True``) without computing
    it.  ``parts``: one tensor ``[B,ny,nx,na,5+C]`` (rows already interleaved) or the head's three
    ``(conf [...,1], cls [...,C], bbox [...,4])``."""

    @staticmethod
    def __new__(cls, parts: Sequence[torch.Tensor], decode: bool):
        if len(parts) == 1:
            base = parts[0].detach()
        else:  # no tensor of the concatenated shape exists: a 4-byte stand-in expanded to that shape carries the metadata
            shape = tuple(parts[1].shape[:-1]) + (sum(int(p.shape[-1]) for p in parts),)
            base = torch.empty(1, dtype=parts[1].dtype, device=parts[1].device).expand(shape)
        return torch.Tensor._make_subclass(cls, base, False)

    def __init__(self, parts: Sequence[torch.Tensor], decode: bool):
        self._bg_parts: Tuple[torch.Tensor, ...] = tuple(parts)
        self._bg_decode = bool(decode)
        self._bg_real: Optional[torch.Tensor] = None

    @property
    def pending(self) -> bool:
        """True while nobody has asked for the values."""
        return self._bg_real is None

    @property
    def parts(self) -> Tuple[torch.Tensor, ...]:
        return self._bg_parts

    @property
    def decode(self) -> bool:
        return self._bg_decode

    def materialize(self) -> torch.Tensor:
        if self._bg_real is None:
            rows = self._bg_parts[0] if len(self._bg_parts) == 1 else torch.cat(self._bg_parts, dim=-1)
            if self._bg_decode:
                from . import ops
                rows = ops.decode_train(rows)
            self._bg_real = rows
        return self._bg_real

    def _meta(self, func, args, kwargs):
        if len(self._bg_parts) == 1:
            return func(*(a._bg_parts[0] if isinstance(a, LazyRows) else a for a in args), **kwargs)
        if func == _T.requires_grad.__get__:
            return any(p.requires_grad for p in self._bg_parts)
        if func == _T.is_contiguous:
            return True
        if func == _T.stride:
            st, acc = [], 1
            for n in reversed(tuple(_T.size(self._bg_parts[1])[:-1]) + (sum(int(p.shape[-1]) for p in self._bg_parts),)):
                st.append(acc)
                acc *= int(n)
            st = tuple(reversed(st))
            return st if len(args) == 1 and not kwargs else st[args[1] if len(args) > 1 else kwargs["dim"]]
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _META and args and isinstance(args[0], LazyRows):
            return args[0]._meta(func, args, kwargs)

        def real(a):
            if isinstance(a, LazyRows):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(real(x) for x in a)
            if isinstance(a, dict):
                return {k: real(v) for k, v in a.items()}
            return a

        return func(*real(args), **real(kwargs))


LazyDecoded = LazyRows  # (name used by the first version of the drop-in)


class HeadTrace(torch.Tensor):
    """Marks the tensors flowing through ``EffiDecHead.forward`` so that its final
    ``torch.cat([conf, cls, bbox], dim=-1)`` (modules/common.py:919) can be recognised and deferred; every other
    function behaves as on a plain tensor."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.cat and args and isinstance(args[0], (list, tuple)) and len(args[0]) == 3:
            parts = args[0]
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            if all(isinstance(p, torch.Tensor) and p.dim() == 5 for p in parts) and dim in (-1, 4) \
                    and int(parts[0].shape[-1]) == 1 and int(parts[2].shape[-1]) == 4 \
                    and all(p.shape[:-1] == parts[0].shape[:-1] for p in parts):
                return LazyRows([p.as_subclass(torch.Tensor) for p in parts], decode=False)
        return super().__torch_function__(func, types, args, kwargs)


def loss_inputs_if_pending(preds):
    """``("raw", [rows x3])`` or ``("split", [(conf, cls, bbox) x3])`` if every element of ``preds`` is a still-pending
    :class:`LazyRows` awaiting the training-mode decode (all of the same kind), else ``None``."""
    if not all(isinstance(p, LazyRows) and p.pending and p.decode for p in preds):
        return None
    kinds = {len(p.parts) for p in preds}
    if kinds == {1}:
        return "raw", [p.parts[0] for p in preds]
    if kinds == {3}:
        return "split", [p.parts for p in preds]
    return None


class LazyPreds(torch.Tensor):
    """Stands for the inference-mode predictions of ``DetectionNet.forward(x, inference=True, og_size=...)``
    (modules/detection.py:69-91) -- one scale ``[B,ny,nx,na,D]``, one scale reshaped ``[B,n,D]``, or the concatenated
    ``[B,N,D]`` -- holding only the head outputs and what the forward would have done to them.  The patched
    ``inference_det.post_process_preds`` recognises the concatenated form and runs the fused decode+NMS on the head
    outputs; any other consumer gets the real tensor (CUDA decode per scale, ``_bbox_to_size``, reshape, cat)."""

    @staticmethod
    def __new__(cls, shape, scales):
        raw = scales[0]["raw"]
        base = torch.empty(1, dtype=raw.dtype, device=raw.device).expand(tuple(shape))
        return torch.Tensor._make_subclass(cls, base, False)

    def __init__(self, shape, scales):
        self._bg_scales = list(scales)   # per scale: raw, anchors, input_shape, rescale (None or (_from, _to)), num_classes
        self._bg_real: Optional[torch.Tensor] = None

    @property
    def pending(self) -> bool:
        return self._bg_real is None

    @property
    def scales(self):
        return self._bg_scales

    def with_rescale(self, _from, _to) -> "LazyPreds":
        return LazyPreds(tuple(_T.size(self)), [dict(sc, rescale=(_from, _to)) for sc in self._bg_scales])

    def materialize(self) -> torch.Tensor:
        if self._bg_real is None:
            from . import ops
            outs = []
            for sc in self._bg_scales:
                d = ops.decode_scale(sc["raw"], sc["anchors"], sc["input_shape"], True, None, sc["num_classes"], sc.get("tanh_cols", 0))
                if sc["rescale"] is not None:
                    d = ops.bbox_to_size(d, sc["rescale"][0], sc["rescale"][1], sc["num_classes"])
                outs.append(d)
            shape = tuple(_T.size(self))
            if len(outs) == 1:
                real = outs[0].reshape(shape)
            else:
                real = torch.cat([o.reshape(shape[0], -1, shape[-1]) for o in outs], dim=1)
            self._bg_real = real
        return self._bg_real

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        me = args[0] if args and isinstance(args[0], LazyPreds) else None
        if me is not None and func in _META:
            if func == _T.is_contiguous:
                return True
            if func == _T.requires_grad.__get__:
                return False
            with torch._C.DisableTorchFunctionSubclass():
                if func == _T.stride:
                    return func(torch.empty(0, device="meta").new_empty(tuple(_T.size(me))), *args[1:], **kwargs)
                return func(*args, **kwargs)
        if me is not None and me.pending:
            shape = tuple(_T.size(me))
            # the shape-only steps of DetectionNet.forward keep the stand-in: reshape(B, -1, D), flatten(1, -2) of [B,N,D]
            if func in (_T.reshape, torch.reshape, _T.view) and len(me._bg_scales) == 1:
                tgt = args[1] if len(args) == 2 and isinstance(args[1], (tuple, list, torch.Size)) else args[1:]
                tgt = tuple(int(v) for v in tgt)
                if len(tgt) == 3 and tgt[0] == shape[0] and tgt[2] == shape[-1] and tgt[1] in (-1, me.numel() // (shape[0] * shape[-1])):
                    return LazyPreds((shape[0], me.numel() // (shape[0] * shape[-1]), shape[-1]), me._bg_scales)
            if func in (_T.flatten, torch.flatten) and len(shape) == 3:
                sd = kwargs.get("start_dim", args[1] if len(args) > 1 else 0)
                ed = kwargs.get("end_dim", args[2] if len(args) > 2 else -1)
                if sd == 1 and ed in (-2, 1):
                    return me
            if func == _T.contiguous:
                return me
        if func is torch.cat and args and isinstance(args[0], (list, tuple)) and len(args[0]) == 3:
            parts = args[0]
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            if dim == 1 and all(isinstance(p, LazyPreds) and p.pending and len(p._bg_scales) == 1 and p.dim() == 3 for p in parts) \
                    and len({(p.shape[0], p.shape[2]) for p in parts}) == 1:
                return LazyPreds((parts[0].shape[0], sum(int(p.shape[1]) for p in parts), parts[0].shape[2]),
                                 [p._bg_scales[0] for p in parts])

        def real(a):
            if isinstance(a, (LazyPreds, LazyRows)):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(real(x) for x in a)
            if isinstance(a, dict):
                return {k: real(v) for k, v in a.items()}
            return a

        return func(*real(args), **real(kwargs))



# ------------------------------------------------------------------------------------------------ f2: inference masks
def _on_gpu(t: torch.Tensor) -> bool:
    return t.is_cuda


class ProtoTrace(torch.Tensor):
    """Marks the prototype tensor ``protos [B, K, Hp, Wp]`` handed to ``inference_seg.post_process_preds`` so that the
    per-image chain of its host loop (inference_seg.py:115-117)

        masks = (coefs @ protos[i].reshape(num_masks, -1)).reshape(-1, *protos.shape[2:]).sigmoid()
        masks = F.interpolate(masks.unsqueeze(dim=0), size=img.shape[1:], mode="bilinear", align_corners=False)
        masks = torch.gt(masks, other=0.5)

    can be recognised: indexing and reshaping keep the mark (the data is the real data), the matrix product with the
    kept rows' coefficients returns a :class:`LazyMasks`.  Every other function behaves as on a plain tensor."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.matmul, _T.matmul, _T.__matmul__, _T.__rmatmul__) and len(args) == 2 and not kwargs:
            a, b = (args[1], args[0]) if func is _T.__rmatmul__ else args
            if isinstance(b, ProtoTrace) and isinstance(a, torch.Tensor) and not isinstance(a, ProtoTrace) \
                    and a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[0] and _on_gpu(a) and _on_gpu(b) \
                    and a.dtype == torch.float32 and b.dtype == torch.float32 and _T.is_contiguous(b) \
                    and not a.requires_grad and not b.requires_grad:
                return LazyMasks(a.as_subclass(torch.Tensor), b.as_subclass(torch.Tensor), ())
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        # only views of the prototypes stay marked (indexing, reshape); anything computed from them is a plain tensor
        if isinstance(out, torch.Tensor) and func in (_T.__getitem__, _T.reshape, torch.reshape, _T.view, _T.contiguous,
                                                      _T.select, torch.select, _T.flatten, torch.flatten):
            return out.as_subclass(ProtoTrace)
        if isinstance(out, torch.Tensor):
            return out.as_subclass(torch.Tensor)
        return out


class LazyMasks(torch.Tensor):
    """Stands for ``coefs @ proto`` (``coefs [n, K]``, ``proto [K, Hp*Wp]`` of one image) and for what
    ``inference_seg.post_process_preds`` does to it next -- reshape to ``[n, Hp, Wp]``, ``sigmoid``, ``unsqueeze(0)``,
    ``F.interpolate(size, "bilinear", align_corners=False)`` -- remembering the steps instead of running them.  The closing
    ``torch.gt(masks, other=0.5)`` runs the two mask kernels (``ops.seg_masks``: one byte per output pixel instead of five
    ATen passes with an ``[n, H, W]`` fp32 intermediate).  Any other step, or any other consumer, gets the reference's
    values: the recorded chain is replayed with ATen first."""

    @staticmethod
    def __new__(cls, coefs, proto, log, shape=None):
        shape = (int(coefs.shape[0]), int(proto.shape[1])) if shape is None else tuple(shape)
        base = torch.empty(1, dtype=coefs.dtype, device=coefs.device).expand(shape)
        return torch.Tensor._make_subclass(cls, base, False)

    def __init__(self, coefs, proto, log, shape=None):
        self._bg_coefs, self._bg_proto, self._bg_log = coefs, proto, tuple(log)
        self._bg_real: Optional[torch.Tensor] = None

    @property
    def pending(self) -> bool:
        return self._bg_real is None

    def materialize(self) -> torch.Tensor:
        if self._bg_real is None:
            x = self._bg_coefs @ self._bg_proto
            for func, rest, kw in self._bg_log:
                x = func(x, *rest, **kw)
            self._bg_real = x
        return self._bg_real

    def _then(self, func, rest, kw) -> "LazyMasks":
        with torch._C.DisableTorchFunctionSubclass():
            shape = func(torch.empty(tuple(_T.size(self)), device="meta"), *rest, **kw).shape
        return LazyMasks(self._bg_coefs, self._bg_proto, self._bg_log + ((func, tuple(rest), dict(kw)),), shape)

    def _fused_size(self):
        """``(Hp, Wp, H, W)`` if the recorded steps are exactly the reference's chain, else ``None``."""
        import torch.nn.functional as F
        log = self._bg_log
        if len(log) != 4 or log[0][0] not in (_T.reshape, torch.reshape, _T.view) or log[1][0] not in (_T.sigmoid, torch.sigmoid) \
                or log[2][0] not in (_T.unsqueeze, torch.unsqueeze) or log[3][0] is not F.interpolate:
            return None
        n, HW = int(self._bg_coefs.shape[0]), int(self._bg_proto.shape[1])
        with torch._C.DisableTorchFunctionSubclass():
            s1 = tuple(log[0][0](torch.empty(n, HW, device="meta"), *log[0][1], **log[0][2]).shape)
        dim = log[2][2].get("dim", log[2][1][0] if log[2][1] else None)
        kw = dict(log[3][2])
        size = kw.pop("size", log[3][1][0] if log[3][1] else None)
        if len(s1) != 3 or s1[0] != n or dim != 0 or len(log[3][1]) > 1 or size is None or len(tuple(size)) != 2 \
                or kw.pop("mode", "nearest") != "bilinear" or kw.pop("align_corners", None) is not False \
                or kw.pop("antialias", False) or any(v is not None for v in kw.values()):
            return None
        return s1[1], s1[2], int(size[0]), int(size[1])

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        import torch.nn.functional as F
        kwargs = kwargs or {}
        me = args[0] if args and isinstance(args[0], LazyMasks) else None
        if me is not None and func in _META:
            if func == _T.is_contiguous:
                return True
            if func == _T.requires_grad.__get__:
                return False
            with torch._C.DisableTorchFunctionSubclass():
                if func == _T.stride:
                    return func(torch.empty(tuple(_T.size(me)), device="meta"), *args[1:], **kwargs)
                return func(*args, **kwargs)
        if me is not None and me.pending and not any(isinstance(a, torch.Tensor) for a in list(args[1:]) + list(kwargs.values())):
            if func in (_T.reshape, torch.reshape, _T.view, _T.sigmoid, torch.sigmoid, _T.unsqueeze, torch.unsqueeze, F.interpolate):
                return me._then(func, args[1:], kwargs)
            if func in (torch.gt, _T.gt, _T.__gt__):
                other = kwargs.get("other", args[1] if len(args) > 1 else None)
                fs = me._fused_size()
                if fs is not None and isinstance(other, float) and other == 0.5 and fs[0] * fs[1] == int(me._bg_proto.shape[1]) \
                        and int(me._bg_coefs.shape[1]) <= 64:
                    from . import ops
                    Hp, Wp, H, W = fs
                    n, K = (int(v) for v in me._bg_coefs.shape)
                    return ops.seg_masks(me._bg_coefs.contiguous(), [n], me._bg_proto.view(1, K, Hp, Wp), (H, W)).unsqueeze(0)

        def real(a):
            if isinstance(a, LazyMasks):
                return a.materialize()
            if isinstance(a, ProtoTrace):
                return a.as_subclass(torch.Tensor)
            if isinstance(a, (list, tuple)):
                return type(a)(real(x) for x in a)
            if isinstance(a, dict):
                return {k: real(v) for k, v in a.items()}
            return a

        return func(*real(args), **real(kwargs))
