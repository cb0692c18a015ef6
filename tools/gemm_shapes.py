#!/usr/bin/env python
"""Per-shape timing of dupl_gemm_bf16x3 (CUDA events, 20 launches each) for the shapes of the hot path.
   python tools/gemm_shapes.py            # prints one line per (shape, epilogue): us, issued TFLOP/s (3 MMAs per product)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dupl_b200 import _lib as L, ops  # noqa: E402


def make_group(M, N, K, epi):
    dev = "cuda"
    bf = dict(dtype=torch.bfloat16, device=dev)
    a = (torch.randn(M, K, device=dev).to(torch.bfloat16), torch.zeros(M, K, **bf))
    w = (torch.randn(N, K, device=dev).to(torch.bfloat16), torch.zeros(N, K, **bf))
    g = dict(a=a, w=w)
    if epi in (L.EPI_F32, L.EPI_RESID):
        g["out_f32"] = torch.empty(M, N, device=dev)
    if epi == L.EPI_RESID:
        g["resid"] = torch.randn(M, N, device=dev)
        g["bias"] = torch.randn(N, device=dev)
    if epi in (L.EPI_SPLIT, L.EPI_GELU_SPLIT):
        g["out"] = (torch.empty(M, N, **bf), torch.empty(M, N, **bf))
        g["bias"] = torch.randn(N, device=dev)
    return g


def bench(M, N, K, epi, name, ksplit=0, reps=20, groups=1):
    dev = "cuda"
    gs = [make_group(M, N, K, epi) for _ in range(groups)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.gemm_bf16x3(gs, M, N, K, epi, ksplit=ksplit)
    ts = []
    for _ in range(reps):
        flush.zero_()                      # evict L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_bf16x3(gs, M, N, K, epi, ksplit=ksplit)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    print(f"{name:34s} M={M:6d} N={N:5d} K={K:5d} groups={groups}  {us:8.1f} us  {groups * 3 * 2.0 * M * N * K / us / 1e6:7.1f} TF/s issued")
    return us


def train_shapes():
    """The 8 GEMMs of one encoder block's backward at M = 3140 token rows (b = 4, 448x448), one student per launch vs both."""
    Mt, Kp = 3140, 3200
    shapes = [("dgrad fc2  [M,3072]<-768", Mt, 3072, 768, L.EPI_F32, 0), ("dgrad fc1  [M,768]<-3072", Mt, 768, 3072, L.EPI_F32, 0),
              ("dgrad proj [M,768]<-768 split", Mt, 768, 768, L.EPI_SPLIT, 0), ("dgrad qkv  [M,768]<-2304", Mt, 768, 2304, L.EPI_F32, 0),
              ("wgrad fc2  [768,3072]", 768, 3072, Kp, L.EPI_F32, 8), ("wgrad fc1  [3072,768]", 3072, 768, Kp, L.EPI_F32, 8),
              ("wgrad proj [768,768]", 768, 768, Kp, L.EPI_F32, 8), ("wgrad qkv  [2304,768]", 2304, 768, Kp, L.EPI_F32, 8)]
    tot = {1: 0.0, 2: 0.0}
    for name, M, N, K, epi, ks in shapes:
        for g in (1, 2):
            tot[g] += bench(M, N, K, epi, name, ksplit=ks, groups=g) * (2 // g)
    print(f"one block, both students: {tot[1]:.1f} us with one student per launch, {tot[2]:.1f} us grouped; x12 blocks = "
          f"{tot[1] * 12 / 1e3:.2f} / {tot[2] * 12 / 1e3:.2f} ms per step")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_shapes()
        sys.exit(0)
    Mc, Mt = 21976, 3140
    for M, tag in ((Mc, "cam"), (Mt, "train")):
        bench(M, 2304, 768, L.EPI_SPLIT, f"{tag} qkv   (split out)")
        bench(M, 768, 768, L.EPI_RESID, f"{tag} proj  (resid out)")
        bench(M, 768, 768, L.EPI_F32, f"{tag} proj-shape, plain f32 out")
        bench(M, 3072, 768, L.EPI_GELU_SPLIT, f"{tag} fc1   (gelu split out)")
        bench(M, 768, 3072, L.EPI_RESID, f"{tag} fc2   (resid out)")
    bench(Mt, 3072, 768, L.EPI_F32, "train dgrad fc2 -> [M,3072]")
    bench(Mt, 768, 2304, L.EPI_F32, "train dgrad qkv -> [M,768]")
    for ks in (0, 8):
        bench(768, 768, 3200, L.EPI_F32, f"train wgrad proj ksplit={ks}", ksplit=ks)
        bench(3072, 768, 3200, L.EPI_F32, f"train wgrad fc1  ksplit={ks}", ksplit=ks)
        bench(2304, 768, 3200, L.EPI_F32, f"train wgrad qkv  ksplit={ks}", ksplit=ks)
