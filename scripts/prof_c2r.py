import sys
sys.path.insert(0, ".")
import torch
from vision_conglomerate_b200 import ops, synth
dev = torch.device("cuda", 0)
raws = [r.to(dev) for r in synth.raw_head_outputs(64, 640, 640, 80, "R", 7)]
anc = [synth.anchors_tensor(s) for s in synth.SCALES]
plan = ops.DetectPlan([tuple(r.shape) for r in raws], anc, (640, 640), 80, dev, None, 0.65, 0.001, 4)
for _ in range(3):
    plan.enqueue(raws); r = plan.result()
print(r.pred_boxes.shape)
