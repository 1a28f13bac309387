"""vision_conglomerate_b200 -- B200 (sm_100a) implementation of the box-geometry hot path of
ches-001/vision-conglomerate: detection-head decode, score/threshold/NMS post-processing, YOLOv5-style
target assignment and the detection loss, behind the reference's own Python call signatures.

    from vision_conglomerate_b200 import ops, dropin
    dropin.install(DetectionDataset, DetectionLoss, DetectionNet)   # train_det.py / inference_det.py unchanged

The arithmetic lives in ``csrc/`` (hand-written CUDA, C ABI in ``include/boxgeom.h``); this package is
the thin ctypes/torch shim over it.  There is no CPU fallback.
"""
from . import synth  # noqa: F401  (pure-Python seeded input generators; no native code needed)

__all__ = ["ops", "dropin", "synth", "_lib"]


def __getattr__(name):
    if name in ("ops", "dropin", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
