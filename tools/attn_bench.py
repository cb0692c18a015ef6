#!/usr/bin/env python
"""Attention forward / backward kernels alone at the training shape (b = 4, N = 785, 12 heads) and the MS-CAM shapes:
back-to-back launches timed with CUDA events (the GPU, not the host, is the limit: 20 calls are queued before the first
finishes), so the per-call figure is device time.  Under `ncu` (tools/gpu_r2_ncu_attn.sh) the same script gives the captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dupl_b200 import ops  # noqa: E402


def run(B, gh, gw, reps=20, backward=True):
    segs, M, _ = ops.make_segments([(B, gh, gw)])
    N = segs[0].tokens
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(M, 2304, generator=g).cuda()
    dO = torch.randn(M, 768, generator=g).cuda()
    q = ops.split_bf16(qkv)
    d = ops.split_bf16(dO)
    oh = torch.zeros(M, 768, dtype=torch.bfloat16, device="cuda")
    ol = torch.zeros_like(oh)
    lse = torch.empty(M, 12, dtype=torch.float32, device="cuda")

    def fwd():
        ops.attention_fwd(q[0], q[1], oh, ol, segs, 12, 0.125, lse=lse)

    def bwd():
        ops.attention_bwd(q, (oh, ol), d, lse, B, N, 12, 0.125)

    out = {}
    for name, fn in (("fwd", fwd),) + ((("bwd", bwd),) if backward else ()):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        flop = B * 12 * 4.0 * N * N * 64 * (1 if name == "fwd" else 2.5)     # fwd: QK^T + PV; bwd: 5 products
        out[name] = us
        print(f"attention {name}  B={B} N={N}: {us:8.1f} us per call, {flop / us / 1e6:6.1f} TFLOP/s algorithmic "
              f"({3 * flop / us / 1e6:6.1f} issued)")
    return out


if __name__ == "__main__":
    reps = int(os.environ.get("REPS", "20"))
    run(4, 28, 28, reps)
    if not os.environ.get("TRAIN_ONLY"):
        run(8, 42, 42, reps, backward=False)
        run(8, 14, 14, reps, backward=False)
