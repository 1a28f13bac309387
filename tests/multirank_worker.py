"""Worker of tests/test_gpu_multirank.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Config-3-shaped step, image-sharded: every rank runs the fused loss forward + backward on its own contiguous image
range, the per-scale terms are combined by ``shard.allreduce_loss_terms`` over NCCL, and the result must equal the
single-GPU big-batch loss (which every rank also computes for the check).  The ranks address their GPU as
``cuda:LOCAL_RANK`` WITHOUT calling ``torch.cuda.set_device`` -- like the reference's DDP path does."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vision_conglomerate_b200 import ops, shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, H, W, C, G = 16, 320, 320, 80, 40
    anc = [synth.anchors_tensor(s) for s in synth.SCALES]
    cfg = dict(synth.LOSS_CONFIG, num_classes=C)
    t_full = synth.targets(B, G, C, 0).to(dev)
    full = [p.to(dev) for p in synth.train_preds(B, H, W, C, 1)]
    out = {"rank": rank}
    for form in ("decoded", "raw"):
        with torch.no_grad():
            big, _ = ops.detection_loss(full, t_full, anc, cfg, with_metrics=False, input_form=form)
        s, e = shard.shard_range(B, world, rank)
        loc = [x[s:e].clone().requires_grad_(True) for x in full]
        tl = shard.shard_targets(t_full, s, e)
        loss, _, sc = ops.detection_loss(loc, tl, anc, cfg, with_metrics=False, return_scalars=True, input_form=form)
        loss.backward()
        cells = [x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] for x in loc]
        comb = shard.allreduce_loss_terms(sc, cells, cfg)
        # gradients: d(big-batch loss)/d(local logits) = local gradient rescaled per term; here only finiteness and
        # the device are checked, the values are covered against the oracle in the single-GPU suite
        ok_grad = all(p.grad is not None and p.grad.device == dev and bool(torch.isfinite(p.grad).all()) for p in loc)
        out[form] = {"big": float(big), "combined": float(comb), "local": float(loss), "grad_ok": ok_grad}
    out["current_device"] = torch.cuda.current_device()
    dist.barrier(device_ids=[local])
    with open(os.path.join(os.environ["BG_MULTIRANK_OUT"], "rank%d.json" % rank), "w") as f:
        json.dump(out, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
