#!/usr/bin/env python
"""Driver M1: run an UNMODIFIED reference script (baseline/_ref/train_final_voc.py, train_final_coco.py)
in this image.

    torchrun --nproc-per-node N baseline/run_script.py [--dropin] [--trace FILE] train_final_voc.py -- <script args>

  --dropin   the script's `from model.model_dupl import siamese_network`, `from model.PAR import PAR`,
             `from model.losses import ...`, `from utils import cam_helper` (and utils.camutils / utils.dcrf) resolve to
             dupl_b200's modules; everything else (datasets, loop body, optimizer, logging) is the reference's own code.
             Without it the stock reference runs (stock torch on the GPU): the comparator.
  --trace F  appends one JSON line per `optim.step()` with the loss parts the script has just computed (read from the
             frame of the script's own loop through sys.settrace-free introspection: the AverageMeter.add call it makes
             every iteration, utils/pyutils.py) so that two runs can be compared iteration by iteration.

The script file itself is executed with runpy from the verbatim copy; not a byte of it is edited.
"""
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    argv = sys.argv[1:]
    dropin, trace = False, None
    while argv and argv[0].startswith("--"):
        if argv[0] == "--dropin":
            dropin = True
            argv = argv[1:]
        elif argv[0] == "--trace":
            trace = argv[1]
            argv = argv[2:]
        else:
            break
    script = argv[0]
    rest = argv[1:]
    if rest and rest[0] == "--":
        rest = rest[1:]
    from baseline import compat
    compat.enter_reference_tree(dropin=dropin)
    path = script if os.path.isabs(script) else os.path.join(compat.REF_ROOT, script)
    if trace:
        import utils.pyutils as pyutils     # the reference's own module (baseline/_ref is first on sys.path)
        orig_add = pyutils.AverageMeter.add
        rank = int(os.environ.get("RANK", "0"))
        count = [0]

        def add(self, d):                   # the script calls avg_meter.add({...loss parts...}) once per iteration
            if rank == 0 and "cls_loss" in d:
                with open(trace, "a") as f:
                    f.write(json.dumps({"iter": count[0], **{k: float(v) for k, v in d.items()}}) + "\n")
                count[0] += 1
            return orig_add(self, d)
        pyutils.AverageMeter.add = add
    os.chdir(compat.REF_ROOT)               # the scripts use relative paths (./pretrained, datasets/voc)
    sys.argv = [path] + rest
    runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    main()
