"""CPU suite for the host side: the C-ABI library loads and exports every declared symbol (no compute
calls without a GPU), the drop-in layer keeps the reference signatures, operators refuse CPU tensors, and
the N>1 image-sharding logic agrees with the single-process result (gloo, world_size 2)."""
import inspect
import os
import re

import numpy as np
import pytest
import torch

from oracle import oracle as O, ref_harness
from vision_conglomerate_b200 import _lib, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "boxgeom.h")).read()
    declared = set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.bg_version() >= 100
    assert L.bg_strerror(0) == b"ok" and b"workspace" in L.bg_strerror(2)
    assert L.bg_sizeof_detect_params() == __import__("ctypes").sizeof(_lib.DetectParams)
    assert L.bg_sizeof_loss_params() == __import__("ctypes").sizeof(_lib.LossParams)
    assert L.bg_sizeof_seg_params() == __import__("ctypes").sizeof(_lib.SegParams)


def test_header_is_plain_c_and_links(tmp_path):
    """include/boxgeom.h is the drop-in boundary: it must compile as strict C99 (no C++-isms, no torch types) and a C
    program must link against libboxgeom.so and call it -- here only entry points that need no GPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "boxgeom.h"\n#include <stdio.h>\n#include <string.h>\n'
        "int main(void) {\n"
        "    bg_detect_params p; bg_loss_params q; bg_seg_params r;\n"
        "    memset(&p, 0, sizeof p); memset(&q, 0, sizeof q); memset(&r, 0, sizeof r);\n"
        "    p.B = 64; p.C = 80; p.na = 3; p.H = 640; p.W = 640;\n"
        "    for (int s = 0; s < 3; ++s) { p.ny[s] = p.nx[s] = 80 >> s; }\n"
        '    printf("%d %s %zu %zu %zu %zu %d\\n", bg_version(), bg_strerror(0), bg_sizeof_detect_params(), sizeof p,\n'
        "           bg_sizeof_loss_params() - sizeof q + bg_sizeof_seg_params() - sizeof r, bg_detect_workspace_bytes(&p, 0),\n"
        "           bg_batched_nms(NULL, NULL, NULL, -1, 0.5, 16, NULL, NULL, NULL, 0, 0, NULL));\n"
        "    return 0;\n}\n")
    libdir = os.path.join(ROOT, "vision_conglomerate_b200", "csrc")
    exe = tmp_path / "abi"
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                         "-o", str(exe), "-L", libdir, "-lboxgeom", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ver, ok, sz_lib, sz_c, diff, ws, rc = out.stdout.split()
    assert int(ver) >= 100 and ok == "ok" and sz_lib == sz_c and int(diff) == 0 and int(ws) > 0 and int(rc) == 1


def test_workspace_queries_run_on_host():
    import ctypes as C
    L = _lib.lib()
    p = _lib.DetectParams()
    p.B, p.C, p.na, p.H, p.W = 64, 80, 3, 640, 640
    for s, v in enumerate((80, 40, 20)):
        p.ny[s] = p.nx[s] = v
    small, big = L.bg_detect_workspace_bytes(C.byref(p), 0), L.bg_detect_workspace_bytes(C.byref(p), 1 << 30)
    assert 0 < small < big and big - small >= (1 << 30)
    p.na = 99
    assert L.bg_detect_workspace_bytes(C.byref(p), 0) == 0  # invalid parameters are rejected, not crashed on
    assert L.bg_batched_nms_workspace_bytes(1000, 16, 0) > 0
    assert L.bg_assign_workspace_bytes(25600, 3) >= (5 * 3 * 25600 // 4096) * 8
    # argument validation happens before any CUDA call
    assert L.bg_batched_nms(None, None, None, -1, 0.5, 16, None, None, None, 0, 0, None) == 1
    assert L.bg_ciou_fwd(None, None, -5, 1e-7, None, None) == 1


def test_operators_refuse_cpu_tensors():
    from vision_conglomerate_b200 import ops
    b, s, i = synth.nms_boxes(10, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.batched_nms(b, s, i, 0.5)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.build_target_by_scale(synth.targets(1, 3), (20, 20), synth.anchors_tensor("lg"))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.detect(synth.raw_head_outputs(1, 64, 64, 3), [synth.anchors_tensor(k) for k in synth.SCALES], (64, 64), 3)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vision_conglomerate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                bad = re.search(r"(from\s+oracle|import\s+oracle|libboxgeom_oracle|oracle[/.](oracle|boxgeom|ref_harness)|bgo_)", src)
                assert not bad, f"{f} references the oracle: {bad.group(0) if bad else ''}"


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present (GPU box)")
def test_dropin_keeps_reference_signatures():
    from vision_conglomerate_b200 import dropin
    import torchvision
    ns = ref_harness.load()
    import inference_seg
    from modules.segmentation_loss import SegmentationLoss
    before_fns = (ns.DetectionNet.forward, ns.EffiDecHead.forward, ns.inference_det.post_process_preds)
    before_seg = (inference_seg.post_process_preds, SegmentationLoss.forward)
    before = {
        "sppp": inspect.signature(inference_seg.post_process_preds),
        "segfwd": inspect.signature(SegmentationLoss.forward),
        "bts": inspect.signature(ns.DetectionDataset.build_target_by_scale),
        "ciou": inspect.signature(ns.DetectionLoss.compute_ciou),
        "fwd": inspect.signature(ns.DetectionLoss.forward),
        "gsp": inspect.signature(ns.DetectionNet._get_scale_pred),
        "b2s": inspect.signature(ns.DetectionNet._bbox_to_size),
        "grid": inspect.signature(ns.DetectionNet._make_2dgrid),
        "rm": inspect.signature(ns.make_anchors.ratio_metrics),
        "rmx": inspect.signature(ns.make_anchors.ratio_metrics_w_extras),
        "nms": inspect.signature(torchvision.ops.batched_nms),
        "netfwd": inspect.signature(ns.DetectionNet.forward),
        "headfwd": inspect.signature(ns.EffiDecHead.forward),
        "ppp": inspect.signature(ns.inference_det.post_process_preds),
    }
    dropin.install(ns.DetectionDataset, ns.DetectionLoss, ns.DetectionNet, make_anchors=ns.make_anchors,
                   EffiDecHead=ns.EffiDecHead, inference_det=ns.inference_det, inference_seg=inference_seg,
                   SegmentationLoss=SegmentationLoss)
    try:
        assert all(dropin.installed().values())
        assert inference_seg.post_process_preds is not before_seg[0] and SegmentationLoss.forward is not before_seg[1]
        after = {
            "sppp": inspect.signature(inference_seg.post_process_preds),
            "segfwd": inspect.signature(SegmentationLoss.forward),
            "bts": inspect.signature(ns.DetectionDataset.build_target_by_scale),
            "ciou": inspect.signature(ns.DetectionLoss.compute_ciou),
            "fwd": inspect.signature(ns.DetectionLoss.forward),
            "gsp": inspect.signature(ns.DetectionNet._get_scale_pred),
            "b2s": inspect.signature(ns.DetectionNet._bbox_to_size),
            "grid": inspect.signature(ns.DetectionNet._make_2dgrid),
            "rm": inspect.signature(ns.make_anchors.ratio_metrics),
            "rmx": inspect.signature(ns.make_anchors.ratio_metrics_w_extras),
            "nms": inspect.signature(torchvision.ops.batched_nms),
            "netfwd": inspect.signature(ns.DetectionNet.forward),
            "headfwd": inspect.signature(ns.EffiDecHead.forward),
            "ppp": inspect.signature(ns.inference_det.post_process_preds),
        }
        for k in before:
            assert list(before[k].parameters) == list(after[k].parameters), k
            for name, prm in before[k].parameters.items():
                assert prm.default == after[k].parameters[name].default or prm.default is inspect._empty, (k, name)
        # the reference's own call sites resolve to the replacement at call time; CPU tensors are refused
        with pytest.raises(RuntimeError, match="CUDA"):
            ns.DetectionDataset.build_target_by_scale(synth.targets(1, 3), (20, 20), synth.anchors_tensor("lg"))
        with pytest.raises(RuntimeError, match="CUDA"):
            torchvision.ops.batched_nms(*synth.nms_boxes(10, 2), 0.5)
        # the segmentation / keypoint variants of build_target_by_scale run on the CUDA kernel too: CPU refused
        with pytest.raises(RuntimeError, match="CUDA"):
            ns.DetectionDataset.build_target_by_scale(synth.targets(2, 3), (20, 20), synth.anchors_tensor("lg"), overlap_masks=False)
        # host tensors keep the reference's own anchor-fit metrics (train_det.py builds them from the label files)
        wh = 0.02 + 0.3 * torch.rand(50, 2)
        anc9 = torch.tensor(sum((synth.ANCHORS[k] for k in synth.SCALES), []))
        assert ns.make_anchors.ratio_metrics(anc9, wh) == _saved_ratio(dropin)(anc9, wh)
        # the cell grid stays the reference's own (cached per shape)
        net = ns.DecodeOnly(80)
        assert torch.equal(ns.DetectionNet._make_2dgrid(net, 5, 4), _saved_grid(dropin)(net, 5, 4))
        # the broadcasting CIoU form (detection_loss.py:231-234) is served by the CUDA operator too: no CPU fallback
        p4 = torch.rand(2, 5, 4) + 0.1
        with pytest.raises(RuntimeError, match="CUDA"):
            ns.DetectionLoss.compute_ciou(p4, p4[:, 0])
    finally:
        dropin.uninstall()
    assert not any(dropin.installed().values())
    assert ns.DetectionNet.forward is before_fns[0] and ns.EffiDecHead.forward is before_fns[1] \
        and ns.inference_det.post_process_preds is before_fns[2]
    assert inference_seg.post_process_preds is before_seg[0] and SegmentationLoss.forward is before_seg[1]
    ref = ns.DetectionDataset.build_target_by_scale(synth.targets(1, 3), (20, 20), synth.anchors_tensor("lg"))
    assert len(ref) == 6


def _saved_ratio(dropin):
    return dropin._saved["ratio_metrics"][1]


def _saved_grid(dropin):
    return dropin._saved["_make_2dgrid"][1]


def test_lazy_decoded_stand_in_mechanics():
    """lazy.LazyDecoded (what the patched _get_scale_pred returns in training mode): metadata comes from the logits
    without materialising; any torch function on it sees the decoded tensor; autograd reaches the logits.  The
    materialiser is stubbed here (the real one is the CUDA decode) with the reference's own arithmetic."""
    from vision_conglomerate_b200 import lazy, ops
    C = 3
    raw = torch.randn(2, 4, 4, 3, 5 + C, requires_grad=True)

    def ref_decode(x):  # modules/detection.py:122,125,164
        xy = x[..., C + 1:C + 3].sigmoid() * 2 - 0.5
        wh = (x[..., C + 3:C + 5].sigmoid() * 2).pow(2)
        return torch.cat((x[..., :C + 1], xy, wh), dim=-1)

    real = ops.decode_train
    ops.decode_train = ref_decode
    try:
        z = lazy.LazyRows([raw], decode=True)
        assert isinstance(z, torch.Tensor) and z.pending
        assert z.shape == raw.shape and z.shape[0] == 2 and z.dim() == 5 and z.dtype == raw.dtype and z.device == raw.device
        assert z.requires_grad and len(z) == 2 and z.numel() == raw.numel() and z.is_contiguous()
        assert z.pending and lazy.loss_inputs_if_pending((z, z, z))[1][0] is raw    # none of that materialised it
        sm, md, lg = (z, z, z)                                                 # tuple packing / unpacking neither
        assert sm.pending
        got = z[..., C + 1:]                                                   # a real consumer: decoded values
        assert not z.pending and lazy.loss_inputs_if_pending((z,)) is None
        assert torch.equal(got, ref_decode(raw)[..., C + 1:]) and type(got) is torch.Tensor
        (z * 1.0).sum().backward()
        exp = torch.autograd.grad(ref_decode(raw).sum(), raw)[0]
        assert torch.allclose(raw.grad, exp)
        z2 = lazy.LazyRows([raw.detach()], decode=True)
        assert torch.equal(torch.cat([z2, z2], 0)[:2], ref_decode(raw.detach()))   # functions taking lists see it too
        # the head's three conv outputs, still apart (EffiDecHead's deferred torch.cat, modules/common.py:919)
        conv = torch.nn.Conv2d(4, 3 * (5 + C), 1)
        feat = torch.randn(2, 4, 4, 4)

        def head(x):  # the shape of EffiDecHead.forward's tail
            y = conv(x)
            pr = lambda t, d: t.permute(0, 2, 3, 1).reshape(2, 4, 4, 3, d)  # noqa: E731
            return torch.cat([pr(y[:, :3], 1), pr(y[:, 3:3 + 3 * C], C), pr(y[:, 3 + 3 * C:], 4)], dim=-1)

        ref_rows = head(feat)
        lz = head(feat.as_subclass(lazy.HeadTrace))
        assert isinstance(lz, lazy.LazyRows) and lz.pending and not lz.decode and len(lz.parts) == 3
        assert lz.shape == ref_rows.shape and lz.requires_grad and lz.is_contiguous() and lz.stride() == ref_rows.stride()
        assert all(type(p) is torch.Tensor for p in lz.parts)
        dec = lazy.LazyRows(lz.parts, decode=True)
        kind, tri = lazy.loss_inputs_if_pending((dec, dec, dec))
        assert kind == "split" and tri[0][1].shape[-1] == C
        assert torch.equal(dec[..., :], ref_decode(ref_rows))            # materialises: cat, then decode
        assert torch.equal(lz + 0, ref_rows)
        conv.zero_grad()
        (dec * 1.0).sum().backward()
        g1 = conv.weight.grad.clone()
        conv.zero_grad()
        ref_decode(head(feat)).sum().backward()
        assert torch.allclose(g1, conv.weight.grad)
        assert lazy.loss_inputs_if_pending((z2, dec)) is None               # mixed kinds are not fused
    finally:
        ops.decode_train = real


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout not present")
def test_lazy_preds_stand_in_mechanics():
    """lazy.LazyPreds (what the patched _get_scale_pred returns with inference=True): the shape-only steps of
    DetectionNet.forward (modules/detection.py:76-91: _bbox_to_size bookkeeping, reshape(B,-1,D), cat(dim=1),
    flatten(1,-2)) keep the stand-in pending; any real consumer gets the reference's tensor.  The CUDA decode is
    stubbed with the reference's own _get_scale_pred / _bbox_to_size."""
    from vision_conglomerate_b200 import lazy, ops
    ns = ref_harness.load()
    C, B, H, W = 5, 2, 64, 96
    net = ns.DecodeOnly(C)
    raws = synth.raw_head_outputs(B, H, W, C, "N", 3)
    anc = [synth.anchors_tensor(k) for k in synth.SCALES]
    real_dec, real_b2s = ops.decode_scale, ops.bbox_to_size
    ops.decode_scale = lambda x, a, ishape, inference, og=None, num_classes=None, tanh_cols=0: net._get_scale_pred(x.clone(), a, input_shape=ishape, inference=inference)
    ops.bbox_to_size = lambda pred, f, t, nc: net._bbox_to_size(pred, f, t)
    try:
        og = (100, 130)
        _from, _to = torch.tensor([W, H, W, H]), torch.tensor([og[1], og[0], og[1], og[0]])
        ref = ref_harness.ref_decode_inference(raws, anc, H, W, og, C)

        def forward_tail(make):       # the tail of DetectionNet.forward, on stand-ins
            ps = [make(r, a) for r, a in zip(raws, anc)]
            ps = [p.with_rescale(_from, _to) for p in ps]
            ps = [p.reshape(B, -1, C + 5) for p in ps]
            return torch.cat(ps, dim=1).flatten(start_dim=1, end_dim=-2)

        mk = lambda r, a: lazy.LazyPreds(tuple(r.shape), [dict(raw=r, anchors=a, input_shape=(H, W), rescale=None, og_size=None, num_classes=C)])  # noqa: E731
        out = forward_tail(mk)
        assert isinstance(out, lazy.LazyPreds) and out.pending and len(out.scales) == 3
        assert out.shape == ref.shape and out.dim() == 3 and out.is_contiguous() and out.stride() == ref.stride() and not out.requires_grad
        assert all(sc["rescale"] is not None for sc in out.scales)
        got = out[..., :]                                    # a real consumer
        assert not out.pending and type(got) is torch.Tensor and torch.equal(got, ref)
        assert torch.equal(torch.sigmoid(forward_tail(mk)[..., :1]), torch.sigmoid(ref[..., :1]))
        one = mk(raws[0], anc[0])
        assert torch.equal(one + 0, net._get_scale_pred(raws[0].clone(), anc[0], input_shape=(H, W), inference=True))
    finally:
        ops.decode_scale, ops.bbox_to_size = real_dec, real_b2s


def test_shard_range_partitions():
    for n, w in [(256, 8), (64, 3), (5, 8), (0, 2)]:
        spans = [shard.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - s for s, e in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, H, W, C, G = 8, 128, 128, 80, 12
        t = synth.targets(B, G, C, 0)
        preds = synth.train_preds(B, H, W, C, 1)
        anc = [synth.anchors_tensor(s) for s in synth.SCALES]
        s, e = shard.shard_range(B, world, rank)
        tl = shard.shard_targets(t, s, e)
        pl = [p[s:e].contiguous() for p in preds]
        cfg = dict(synth.LOSS_CONFIG, num_classes=C)
        rows, cells = [], []
        for p, a, w in zip(pl, anc, cfg["scale_w"]):   # per-rank scalars, exactly what bg_loss_fwd emits
            scal, _, _, _ = O.loss_scale(p, tl, a, cfg, w)
            rows.append(torch.tensor(scal))
            cells.append(p.shape[0] * p.shape[1] * p.shape[2] * p.shape[3])
        loss = shard.allreduce_loss_terms(torch.stack(rows), cells, cfg)
        # image-sharded assignment == the matching slice of the single-process assignment
        idx, cls, _, box = O.build_target_by_scale(tl, (16, 16), anc[0])
        full = O.build_target_by_scale(t, (16, 16), anc[0])
        sel = (full[0][0] >= s) & (full[0][0] < e)
        same = (np.array_equal(np.sort(idx[0] + s), np.sort(full[0][0][sel])) and
                np.array_equal(np.sort(box, axis=0), np.sort(full[3][sel], axis=0)))
        q.put((rank, float(loss), bool(same)))
    finally:
        dist.destroy_process_group()


def test_image_sharded_loss_matches_big_batch_gloo():
    import torch.multiprocessing as mp
    world, port = 2, 29000 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, H, W, C, G = 8, 128, 128, 80, 12
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ref, _, _, _ = O.detection_loss(synth.train_preds(B, H, W, C, 1), synth.targets(B, G, C, 0), anc, synth.LOSS_CONFIG)
    for rank, loss, same in res:
        assert same, "sharded assignment differs"
        assert abs(loss - ref) <= 1e-9 * abs(ref), (rank, loss, ref)


def test_lazy_masks_recognise_the_reference_chain(monkeypatch):
    """lazy.ProtoTrace / lazy.LazyMasks (inference_seg.py:115-117): the reference's chain
    ``sigmoid((coefs @ protos[i].reshape(K, -1)).reshape(-1, Hp, Wp))`` -> ``F.interpolate`` -> ``torch.gt(0.5)`` reaches
    ops.seg_masks with the image's prototypes and the output size, nothing is computed before; any other chain or
    consumer sees the values ATen would have produced.  (Host logic only: the kernel call is replaced by a recorder.)"""
    import torch
    import torch.nn.functional as F
    from vision_conglomerate_b200 import lazy, ops
    calls = []

    def fake(coefs, counts, protos, size):
        calls.append((tuple(coefs.shape), list(counts), tuple(protos.shape), tuple(size)))
        m = (coefs @ protos[0].reshape(protos.shape[1], -1)).reshape(-1, *protos.shape[2:]).sigmoid()
        return F.interpolate(m.unsqueeze(0), size=size, mode="bilinear", align_corners=False)[0] > 0.5

    monkeypatch.setattr(lazy, "_on_gpu", lambda t: True)
    monkeypatch.setattr(ops, "seg_masks", fake)
    g = torch.Generator().manual_seed(0)
    protos, coefs, i = torch.randn(3, 8, 6, 10, generator=g), torch.randn(5, 8, generator=g), torch.tensor(1)
    p = protos.as_subclass(lazy.ProtoTrace)

    def chain(pr, mode="bilinear"):
        m = (coefs @ pr[i].reshape(pr.shape[1], -1)).reshape(-1, *pr.shape[2:]).sigmoid()
        kw = dict(align_corners=False) if mode == "bilinear" else {}
        return F.interpolate(m.unsqueeze(dim=0), size=torch.Size([24, 30]), mode=mode, **kw)

    lz = chain(p)
    assert isinstance(lz, lazy.LazyMasks) and lz.pending and tuple(lz.shape) == (1, 5, 24, 30) and not calls
    out = torch.gt(lz, other=0.5).squeeze(dim=0).detach().cpu().numpy()
    assert calls == [((5, 8), [5], (1, 8, 6, 10), (24, 30))]
    assert (out == torch.gt(chain(protos), other=0.5)[0].numpy()).all()
    # another consumer / another chain: the reference's values, no kernel call
    assert torch.equal(chain(p).sum(), chain(protos).sum())
    assert torch.equal(chain(p, "nearest") > 0.5, chain(protos, "nearest") > 0.5) and len(calls) == 1
    assert torch.equal(torch.gt(chain(p), other=0.25), torch.gt(chain(protos), other=0.25)) and len(calls) == 1
    assert type(p[0]) is lazy.ProtoTrace and type(p * 2) is torch.Tensor and type(p.sum()) is torch.Tensor
