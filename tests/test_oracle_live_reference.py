"""The oracle against the UNMODIFIED reference, live, on seeds that are not in tests/golden/ (-m "not gpu"; skipped
where no reference tree is present).  The committed fixtures pin the oracle to outputs of the reference generated once;
this suite re-derives the same agreement from the reference's own code every run: assignment indices bit- and
order-exact, keep-lists exact as sets (torchvision's order inside score ties is arbitrary), decoded boxes / loss /
gradients within the path's fp32 tolerances (rtol 1e-5, gradients 1e-4)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O, ref_harness
from vision_conglomerate_b200 import synth
from tests.util import assert_close

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ns():
    n = torch.get_num_threads()
    torch.set_num_threads(1)   # (the reference's duplicate-index scatter is racy with more threads)
    yield ref_harness.load()
    torch.set_num_threads(n)


@pytest.mark.parametrize("seed,B,G,S", [(101, 3, 7, 160), (202, 2, 20, 256), (303, 5, 1, 96)])
def test_assignment_live(ns, seed, B, G, S):
    """DetectionDataset.build_target_by_scale (dataset/detection_dataset.py:90-246) vs oracle: every output, every scale."""
    t = synth.targets(B, G, 80, seed, fixed=False)
    for (ny, nx), sc in zip(synth.fmap_shapes(S, S), synth.SCALES):
        a = synth.anchors_tensor(sc)
        ri, rc, ra, rb, _, _ = ns.DetectionDataset.build_target_by_scale(t.clone(), (ny, nx), a.clone())
        oi, oc, oa, ob = O.build_target_by_scale(t.numpy(), (ny, nx), a.numpy())
        assert all(np.array_equal(x.numpy(), y) for x, y in zip(ri, oi)), (seed, sc)
        assert np.array_equal(rc.numpy(), oc)
        assert np.array_equal(ra.numpy().view(np.uint32), oa.view(np.uint32))      # fp32 outputs are bit-reproducible too
        assert np.array_equal(rb.numpy().view(np.uint32), ob.view(np.uint32))


@pytest.mark.parametrize("seed,dist,iou,thr,tracked", [(11, "T", 0.65, 0.001, None), (12, "TP", 0.35, 0.3, (1, 4, 7, 16, 17)),
                                                       (13, "N", 0.5, 0.2, None)])
def test_decode_and_post_process_live(ns, seed, dist, iou, thr, tracked):
    """_get_scale_pred x3 + cat (modules/detection.py:69-91) and post_process_preds lines 57-97,107-109 vs oracle."""
    B, S, C = 2, 96, 80
    raws = synth.raw_head_outputs(B, S, S, C, dist, seed)
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    with torch.no_grad():
        pr = ref_harness.ref_decode_inference(raws, anc, S, S, (120, 200), C)
    po = O.decode_inference(raws, anc, S, S, (120, 200))
    assert_close(po, pr.numpy(), rtol=1e-5, atol=2e-5 * 200, what="decoded preds")
    # same decoded tensor into both post-processing paths: keep-lists must be identical sets
    cap = ref_harness.ref_post_process(pr.clone(), C, iou, thr, 4, list(tracked) if tracked else None)
    ref_keep = cap["keep"].numpy()
    ref_keep = ref_keep[cap["scores"].numpy()[ref_keep] > np.float32(thr)]
    out = O.post_process(pr.numpy(), iou, thr, 4, tracked)
    keep_o = out["keep"]
    if tracked:   # the reference filters classes per image after the threshold; rebuild that from its captured rows
        n_ref = sum(len(r) for r in cap["per_image"])
        assert n_ref == len(keep_o)
        rows_r = np.concatenate(cap["per_image"]) if cap["per_image"] else np.zeros((0, 6), np.float32)
        assert_close(np.sort(rows_r[:, 0]), np.sort(out["pred_boxes"][:, 0]), rtol=1e-6, atol=0, what="scores of the kept rows")
    else:
        assert np.array_equal(np.sort(ref_keep), np.sort(keep_o))
    assert len(keep_o) > 0 or dist == "T"


@pytest.mark.parametrize("seed,B,G,S,C", [(7, 2, 6, 128, 80), (8, 3, 2, 96, 5)])
def test_loss_live(ns, seed, B, G, S, C):
    """DetectionLoss.forward + backward (modules/detection_loss.py:84-226) vs oracle, decoded form."""
    t = synth.targets(B, G, C, seed, fixed=False)
    g = torch.Generator().manual_seed(seed)
    preds = [torch.randn(B, ny, nx, 3, 5 + C, generator=g) for ny, nx in synth.fmap_shapes(S, S)]
    anchors = {k: synth.ANCHORS[k] for k in synth.SCALES}
    mod = ns.DetectionLoss(ns.FakeModel(C, anchors), **synth.LOSS_CONFIG)
    leaves = [p.clone().requires_grad_(True) for p in preds]
    loss, met = mod(tuple(leaves), t.clone())
    loss.backward()
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    ol, om, og, _ = O.detection_loss([p.numpy() for p in preds], t.numpy(), [a.numpy() for a in anc], dict(synth.LOSS_CONFIG, num_classes=C),
                                     with_grad=True)
    assert_close(ol, float(loss.detach()), rtol=1e-5, atol=0, what="loss")
    for k in ("mean_ciou", "conf_loss", "class_loss", "avg_pos_conf", "avg_neg_conf"):
        assert_close(om[k], met[k], rtol=2e-5, atol=1e-7, what=k)
    for a, b in zip(og, leaves):
        assert_close(a, b.grad.numpy(), rtol=1e-4, atol=1e-8, what="grad preds")
