"""Deferred training-mode decode: how the fused loss reaches the head's logits with zero edits to the host code.

In the reference, ``DetectionNet.forward`` (modules/detection.py:58-96) passes each head output through
``_get_scale_pred(inference=False)`` -- ``xy = 2*sigmoid - 0.5``, ``wh = (2*sigmoid)**2`` (:122,125) over ~15 ATen
kernels that copy the whole ``[B,ny,nx,na,5+C]`` tensor -- and the trainer hands the three results straight to
``DetectionLoss.forward`` (pipeline/detection_trainer.py:178-180).  The CUDA loss applies that decode in registers
(``BG_LOSS_RAW``), so the decoded tensors never need to exist.

The patched ``_get_scale_pred`` therefore returns a :class:`LazyDecoded`: a ``torch.Tensor`` subclass that *stands
for* the decoded tensor (same shape, dtype, device, ``requires_grad``) but only remembers the logits.  The patched
``DetectionLoss.forward`` recognises it and feeds the logits to the fused loss.  Any *other* use -- indexing,
arithmetic, ``torch.cat``, printing, a different loss -- goes through ``__torch_function__``, which materialises the
decoded tensor first (once, with the differentiable CUDA decode ``ops.decode_train``) and then runs the requested
function on it: the values every other consumer sees are exactly the reference's.
"""
from __future__ import annotations

import torch

_T = torch.Tensor
# metadata that the decoded tensor shares with the logits: answered without materialising
_META = {
    _T.shape.__get__, _T.device.__get__, _T.dtype.__get__, _T.ndim.__get__, _T.is_cuda.__get__, _T.layout.__get__,
    _T.requires_grad.__get__, _T.size, _T.dim, _T.ndimension, _T.numel, _T.nelement, _T.stride, _T.is_contiguous,
    _T.__len__, _T.is_floating_point, _T.is_complex, _T.get_device, _T.element_size,
}


class LazyDecoded(torch.Tensor):
    """Stands for ``DetectionNet._get_scale_pred(raw, ..., inference=False)`` without computing it."""

    @staticmethod
    def __new__(cls, raw: torch.Tensor):
        return torch.Tensor._make_subclass(cls, raw.detach(), False)

    def __init__(self, raw: torch.Tensor):
        self._bg_raw = raw
        self._bg_dec = None

    @property
    def pending(self) -> bool:
        """True while nobody has asked for the decoded values."""
        return self._bg_dec is None

    @property
    def logits(self) -> torch.Tensor:
        return self._bg_raw

    def materialize(self) -> torch.Tensor:
        if self._bg_dec is None:
            from . import ops
            self._bg_dec = ops.decode_train(self._bg_raw)
        return self._bg_dec

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _META:
            return func(*(a._bg_raw if isinstance(a, LazyDecoded) else a for a in args), **kwargs)

        def real(a):
            if isinstance(a, LazyDecoded):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(real(x) for x in a)
            if isinstance(a, dict):
                return {k: real(v) for k, v in a.items()}
            return a

        return func(*real(args), **real(kwargs))


def logits_if_pending(preds):
    """The three logit tensors if every element of ``preds`` is a still-pending :class:`LazyDecoded`, else None."""
    if all(isinstance(p, LazyDecoded) and p.pending for p in preds):
        return [p.logits for p in preds]
    return None
