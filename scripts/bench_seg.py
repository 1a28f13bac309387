#!/usr/bin/env python
"""SegmentationLoss forward + backward: the CUDA path (fused detection terms + mask-term kernels) against the
unmodified reference on torch-CUDA on the same GPU (baseline/_ref), same seeded inputs.  Prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vision_conglomerate_b200 import _lib, ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
cfg = dict(synth.LOSS_CONFIG, seg_w=1.0)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


QUICK = "--quick" in sys.argv     # one case, no reference run (for ncu)
MASKS_ONLY = "--masks-only" in sys.argv
for B, S, C, K, G in (() if MASKS_ONLY else ((16, 640, 80, 32, 8),) if QUICK else ((16, 640, 80, 32, 8), (64, 640, 80, 32, 20))):
    preds, protos, t, masks = synth.seg_inputs(B, S, S, C, K, G, seed=11)
    preds = [p.to(dev).requires_grad_(True) for p in preds]
    protos = protos.to(dev).requires_grad_(True)
    t, masks = t.to(dev), masks.to(dev)

    def ours():
        for p in preds:
            p.grad = None
        protos.grad = None
        loss, _ = ops.segmentation_loss(preds, t, protos, masks, anc, cfg, C, K, with_metrics=False)
        loss.backward()
        return loss

    l0 = _lib.launch_count()
    loss = ours()
    launches = _lib.launch_count() - l0
    ms = timed(ours, 20)
    out = {"case": "SegmentationLoss fwd+bwd B=%d %dx%d C=%d K=%d, <=%d gt/img, protos %dx%d" % (B, S, S, C, K, G, S // 2, S // 2),
           "ours_ms": ms, "ours_img_s": B / ms * 1e3, "loss": float(loss), "kernels_per_step": int(launches)}
    try:
        if QUICK:
            raise RuntimeError("skipped (--quick)")
        from oracle import ref_harness
        ns = ref_harness.load()
        mod = ns.SegmentationLoss(ns.FakeSegModel(C, synth.ANCHORS, K).to(dev), overlap_masks=True, **cfg)

        def ref():
            for p in preds:
                p.grad = None
            protos.grad = None
            loss, _ = mod(tuple(preds), t, protos, masks)
            loss.backward()
            return loss

        rl = float(ref())
        rms = timed(ref, 3)
        out.update(reference_torch_cuda_ms=rms, reference_loss=rl, speedup=rms / ms)
    except Exception as e:  # noqa: BLE001
        out["reference"] = repr(e)[:200]
    print(json.dumps(out), flush=True)

# inference side (inference_seg.py:115-117): masks of the kept rows, ours vs the same torch calls per image
B, K, Hp, Wp, H, W = 8, 32, 160, 160, 640, 640
g = torch.Generator().manual_seed(5)
counts = torch.full((B,), 25)
coefs = torch.tanh(torch.randn(int(counts.sum()), K, generator=g)).to(dev)
protos = torch.randn(B, K, Hp, Wp, generator=g).to(dev)


def ours_masks():
    return ops.seg_masks(coefs, counts, protos, (H, W))


def torch_masks():
    r, outs = 0, []
    for i, c in enumerate(counts.tolist()):
        m = (coefs[r:r + c] @ protos[i].reshape(K, -1)).reshape(-1, Hp, Wp).sigmoid()
        m = torch.nn.functional.interpolate(m.unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False)
        outs.append(torch.gt(m, 0.5).squeeze(0))
        r += c
    return outs


if not QUICK:
    a, b = timed(ours_masks, 20), timed(torch_masks, 20)
    print(json.dumps({"case": "masks of 200 kept rows (8 images x 25), K=32, protos 160x160 -> 640x640 bool", "ours_ms": a,
                      "torch_cuda_ms": b, "speedup": b / a, "output_MB": 200 * H * W / 1e6}), flush=True)

if not QUICK:   # the two kernels alone (no Python, no offsets copy)
    from vision_conglomerate_b200.ops import _stream
    L = _lib.lib()
    n = int(counts.sum())
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(counts, 0)
    off = off.to(dev)
    low = torch.empty(n, Hp * Wp, device=dev)
    out = torch.empty(n, H, W, dtype=torch.uint8, device=dev)
    kms = timed(lambda: L.bg_seg_masks(coefs.data_ptr(), off.data_ptr(), protos.data_ptr(), B, K, Hp, Wp, n, H, W, low.data_ptr(),
                                       out.data_ptr(), _stream(dev)), 50)
    byts = n * H * W + 2 * n * Hp * Wp * 4 + B * K * Hp * Wp * 4
    print(json.dumps({"case": "bg_seg_masks kernels only", "ms": kms, "algorithmic_MB": byts / 1e6, "GB_per_s": byts / kms / 1e6}), flush=True)
