#!/usr/bin/env python
"""On-GPU evidence for the tensor-core operand scheme: the 448x448 parity measurements of
tests/fullsize_checks.py (MS-CAM error in un-normalised CAM units, refine-label mismatch, loss parts, gradient errors of
all 308 parameter tensors) for

    bf16x3      the product library: operands as two bf16 planes, 3 MMAs per product (~16 mantissa bits);
    tf32_all    experiment build, every operand rounded (RN) to TF32 before the split = single-pass kind::tf32 everywhere
                (2 bf16-MMA-equivalents per product instead of 3: the most a cheaper scheme can save is 1/3 of the MMA time);
    tf32_lin    experiment build, TF32 operands for the linear layers / convs, the attention kernels' internal operands
                (P, dS, O) at full split precision.

`make -C dupl_b200/csrc emu` builds the experiment libraries; each variant runs in its own process (DUPL_LIB).  The CPU
oracle results are cached, so only the first variant pays for them.

    python tools/precision_table.py [--out gpurun_out/precision_table.json]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {"bf16x3": "libdupl.so", "tf32_all": "libdupl_emu_tf32_all.so", "tf32_lin": "libdupl_emu_tf32_lin.so"}

WORKER = r"""
import json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import fullsize_checks as FC
out = {}
out["attention_785"] = FC.attention_errors(2, 28, 28, seed=28)
out["attention_1765"] = FC.attention_errors(1, 42, 42, seed=42)
m = FC.build_model(21, -3)[0]
out["mscam448"] = FC.mscam_errors(m, 21, -3, b=1, S=448, seed=3, pair=True)
del m
out["train448"] = FC.train_step_errors(21, b=1, S=448, seed=40, n_iter=3000)
print("RESULT " + json.dumps(out))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "precision_table.json"))
    ap.add_argument("--variants", default=",".join(VARIANTS))
    args = ap.parse_args()
    results = {}
    for name in args.variants.split(","):
        lib = os.path.join(ROOT, "dupl_b200", VARIANTS[name])
        if not os.path.exists(lib):
            results[name] = {"error": f"{lib} not built (make -C dupl_b200/csrc emu)"}
            continue
        env = dict(os.environ, DUPL_LIB=lib)
        p = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=3000)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        results[name] = json.loads(line[-1][7:]) if line else {"error": p.stderr[-1500:]}
        print(name, json.dumps(results[name])[:400], flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    rows = [("attention fwd N=785 (max-rel)", "attention_785", "fwd"), ("attention dQ N=785", "attention_785", "dq"),
            ("attention dK N=785", "attention_785", "dk"), ("attention fwd N=1765", "attention_1765", "fwd"),
            ("MS-CAM student 1 (un-normalised CAM units)", "mscam448", "cam1"), ("MS-CAM aux student 1", "mscam448", "aux1"),
            ("MS-CAM student 2", "mscam448", "cam2"), ("refine label mismatch, student 1", "mscam448", "label1_mismatch"),
            ("refine label mismatch, student 2", "mscam448", "label2_mismatch"),
            ("total loss (rel)", "train448", "loss"), ("cls loss", "train448", "cls_loss"), ("ptc loss", "train448", "ptc_loss"),
            ("seg loss", "train448", "seg_loss"), ("label mismatch in the step", "train448", "label_mismatch"),
            ("gradient, worst of 308 tensors (norm-rel)", "train448", "grad_worst"),
            ("gradient, median", "train448", "grad_median"), ("gradients >= 1e-3", "train448", "grad_over_1e-3")]
    names = [n for n in args.variants.split(",")]
    md = ["| quantity at 448x448 (bar: 1e-3) | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    for title, grp, key in rows:
        cells = []
        for n in names:
            v = results.get(n, {}).get(grp, {}).get(key)
            cells.append("n/a" if v is None else (f"{v:.2e}" if isinstance(v, float) else str(v)))
        md.append(f"| {title} | " + " | ".join(cells) + " |")
    with open(os.path.splitext(args.out)[0] + ".md", "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
