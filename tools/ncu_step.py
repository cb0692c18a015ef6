#!/usr/bin/env python
"""One replay of the captured training step (BASELINE metric: VOC phase B, b = 4, 448x448) inside a cudaProfilerStart/Stop
range, for `ncu --profile-from-start off` (launch list and --set full captures: tools/gpu_r2_ncu.sh) and, with --table, a
torch.profiler kernel table of the same replay (no ncu needed)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images  # noqa: E402
from dupl_b200.model.model_dupl import siamese_network  # noqa: E402
from dupl_b200.train_step import Args, TrainStep, make_optimizer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--phase", default="B", choices=["B", "C"])
ap.add_argument("--table", action="store_true")
ap.add_argument("--gaps", action="store_true", help="with --table: idle time between consecutive kernels of the replay, by the kernel that follows the gap")
a = ap.parse_args()
m = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
m.load_state_dict(init_state_dict(21), strict=True)
m = m.cuda().train()
step = TrainStep(m, make_optimizer(m, capturable=True), args=Args, capture=True)
x = synth_images(4, 448, 448, seed=0).cuda()
cls, box = synth_cls_labels(4, 20, seed=0).cuda(), synth_boxes(4, 448, 448, seed=0)
n0 = 9000 if a.phase == "C" else 5000
for i in range(3):
    step(x, cls, box, n0 + i)
torch.cuda.synchronize()
if a.table:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(x, cls, box, n0 + 5)
        if a.gaps:
            step(x, cls, box, n0 + 6)      # two replays: the idle time BETWEEN two steps shows up as well
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=72))
    if a.gaps:
        import collections
        ev = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                     if getattr(e, "device_type", None) is not None and "cuda" in str(e.device_type).lower()), key=lambda t: t[0])
        span = ev[-1][1] - ev[0][0]
        busy, gaps, by_next, hist = 0.0, 0.0, collections.defaultdict(lambda: [0.0, 0]), collections.Counter()
        end = ev[0][0]
        for s0, s1, name in ev:
            g = max(0.0, s0 - end)
            gaps += g
            key = name.split("(")[0].replace("void ", "").replace("dupl::", "")[:48]
            by_next[key][0] += g
            by_next[key][1] += 1
            hist[min(int(g), 20)] += 1
            busy += max(0.0, s1 - max(s0, end))
            end = max(end, s1)
        t0 = ev[0][0]
        end, prev = ev[0][0], "-"
        print("\nidle periods > 3 us (offset ms: previous activity -> idle us -> next activity):")
        for s0, s1, name in ev:
            if s0 - end > 3.0:
                print(f"  {(s0 - t0) / 1e3:8.3f}: {prev[:56]:56s} -> {s0 - end:8.1f} -> {name[:70]}")
            if s1 >= end:
                end, prev = s1, name
        print(f"\nGAPS: {len(ev)} device activities, span {span / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms, idle between activities {gaps / 1e3:.3f} ms")
        print("gap histogram (us -> count):", dict(sorted(hist.items())))
        for k, (g, n) in sorted(by_next.items(), key=lambda kv: -kv[1][0])[:25]:
            print(f"  {g:9.1f} us before {n:4d} x {k}  ({g / n:.2f} us each)")
else:
    torch.cuda.profiler.start()
    step(x, cls, box, n0 + 5)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
