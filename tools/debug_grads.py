#!/usr/bin/env python
"""Debug aid: per-loss-part gradient comparison CUDA vs oracle at 448x448 (which term / which student is off)."""
import json
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fullsize_checks as FC  # noqa: E402
from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images  # noqa: E402


def main(num_classes=81, seed=48, n_iter=20000, coco=True, S=448, b=1):
    from dupl_b200.train_step import Args, CocoArgs, TrainStep
    from oracle import dupl_oracle as O
    P = init_state_dict(num_classes)
    x = synth_images(b, S, S, seed=seed)
    cls = synth_cls_labels(b, num_classes - 1, seed=seed + 1)
    if coco:
        cls = cls.to(torch.uint8)
    box = synth_boxes(b, S, S, seed=seed + 2)
    names = [k for k in P if "pos_embed" not in k and ".head." not in k]
    Pg = {k: v.clone().requires_grad_(k in names) for k, v in P.items()}
    cfg = O.COCO_CFG if coco else O.VOC_CFG
    args = CocoArgs if coco else Args
    loss, parts, labels = O.train_losses(Pg, x, cls, box, n_iter, cfg, thres_target=list(args.high_thres_target))
    m, _ = FC.build_model(num_classes, aux_layer=9 if coco else -3, train=True)
    step = TrainStep(m, None, args=args)
    print("labels: fg px per student (oracle)", [int(((l != 0) & (l != 255)).sum()) for l in labels], "bg", [int((l == 0).sum()) for l in labels])
    for part in ("cls_loss", "ptc_loss", "seg_loss", "sim_loss"):
        ref = torch.autograd.grad(parts[part], [Pg[k] for k in names], retain_graph=True, allow_unused=True)
        m.zero_grad(set_to_none=True)
        gl, gparts, glabels = step.losses(x.cuda(), cls.cuda(), box, n_iter)
        gparts[part].backward()
        torch.cuda.synchronize()
        gp = dict(m.named_parameters())
        rows = []
        for k, r in zip(names, ref):
            g = gp[k].grad
            if r is None or r.abs().max() == 0:
                if g is not None and g.abs().max() > 0:
                    rows.append((float("inf"), k))
                continue
            if g is None:
                rows.append((float("inf"), k + " (missing)"))
                continue
            rows.append((FC.nrel(g, r), k))
        rows.sort(reverse=True)
        per_branch = {br: max([e for e, k in rows if k.startswith(br)] or [0.0]) for br in ("branch1", "branch2")}
        print(part, "value", float(gparts[part]), float(parts[part]), "worst per branch", per_branch)
        print("   top:", [(round(e, 5), k) for e, k in rows[:6]])
        print("   bottom:", [(round(e, 6), k) for e, k in rows[-3:]])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "voc":
        main(21, 40, 3000, False)
    else:
        main()
