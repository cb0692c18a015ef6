#!/usr/bin/env python
"""DenseCRF mean-field: GPU (dupl_b200.utils.dcrf, crf.cu) vs the CPU C restatement (oracle/densecrf_ref.c,
"restatement, not pydensecrf") at the VOC (500x375x21) and COCO (640x480x81) sizes of SURVEY §8(d).
Prints one JSON line per configuration.   python tools/bench_crf.py [--iters 10] [--reps 10] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(H, W, C, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1).float()[None]
    img = torch.nn.functional.avg_pool2d(t, 5, 1, 2, count_include_pad=False)[0].permute(1, 2, 0).round().numpy().astype(np.uint8)
    lg = rng.randn(C, H // 16 + 1, W // 16 + 1).astype(np.float32) * 2.0
    lg = torch.nn.functional.interpolate(torch.from_numpy(lg)[None], size=(H, W), mode="bilinear", align_corners=False)[0]
    return np.ascontiguousarray(img), torch.softmax(lg, 0).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from dupl_b200 import ops
    from oracle.densecrf_ref import DenseCRF as RefCRF
    params = dict(pos_w=1.0, pos_xy_std=1.0, bi_w=4.0, bi_xy_std=121.0, bi_rgb_std=5.0)  # tools/eval_seg_voc.py:104-111
    for name, (H, W, C) in {"voc_500x375x21": (375, 500, 21), "coco_640x480x81": (480, 640, 81)}.items():
        img, p = synth(H, W, C, seed=C)
        img_d, p_d = torch.from_numpy(img).cuda(), torch.from_numpy(p).cuda()
        ws = ops.CrfWorkspace()
        for _ in range(3):
            q, sizes = ops.crf_inference(img_d, p_d, args.iters, ws=ws, **params)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            q, sizes = ops.crf_inference(img_d, p_d, args.iters, ws=ws, **params)
        e1.record()
        torch.cuda.synchronize()
        gpu_ms = e0.elapsed_time(e1) / args.reps
        # end to end like the reference's call: numpy in, numpy out
        from dupl_b200.utils.dcrf import DenseCRF
        crf = DenseCRF(args.iters, 1, 1, 4, 121, 5)
        crf(img, p)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            out = crf(img, p)
        e2e_ms = (time.perf_counter() - t0) * 1000 / args.reps
        line = {"metric": "densecrf_ms_per_image", "config": name, "iters": args.iters, "gpu_ms": gpu_ms, "gpu_e2e_numpy_ms": e2e_ms,
                "lattice_vertices": {"gaussian_d2": sizes[0], "bilateral_d5": sizes[1]}, "pixels": H * W, "classes": C}
        if not args.no_cpu:
            ref = RefCRF(args.iters, 1, 1, 4, 121, 5)
            t0 = time.perf_counter()
            want = ref(img, p)
            cpu_s = time.perf_counter() - t0
            line.update({"cpu_ms": cpu_s * 1000, "cpu_kind": "C restatement (not pydensecrf), 1 thread", "speedup_vs_cpu_1thread": cpu_s * 1000 / gpu_ms,
                         "max_abs_diff_vs_cpu": float(np.abs(out - want).max()), "label_agreement": float((out.argmax(0) == want.argmax(0)).mean())})
        print(json.dumps(line))


if __name__ == "__main__":
    main()
