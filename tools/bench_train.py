#!/usr/bin/env python
"""Full dual-student training step (phase B of train_final_voc.py: MS-CAM + PAR pseudo-labels + both students'
forward/backward + all losses + AdamW) on synthetic VOC batches, b=4 per GPU, 448x448 — BASELINE.json configs[2].

    python tools/bench_train.py [--steps K] [--warmup W] [--batch 4]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...   (DDP over NCCL)

Prints one JSON line on rank 0 (images/s aggregated over ranks, max-over-ranks device time)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--profile", action="store_true", help="after timing, one extra step under torch.profiler (kernel table to stderr)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images
    from dupl_b200.model.model_dupl import siamese_network
    from dupl_b200.train_step import PhaseBStep, make_optimizer

    model = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    model.load_state_dict(init_state_dict(21), strict=True)
    model = model.to(dev).train()
    optim = make_optimizer(model)
    wrapped = model
    if world > 1:
        wrapped = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], find_unused_parameters=True)
    step = PhaseBStep(wrapped, optim, device=dev, reuse_forward=not os.environ.get("DUPL_NO_REUSE"))
    x = synth_images(args.batch, args.size, args.size, seed=rank).to(dev)
    cls = synth_cls_labels(args.batch, 20, seed=rank).to(dev)
    box = synth_boxes(args.batch, args.size, args.size, seed=rank)

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for i in range(args.warmup):
        loss, _ = step(x, cls, box, 3000 + i)
        losses.append(loss.item())
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss, parts = step(x, cls, box, 3000 + args.warmup + i)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "train_images_per_sec_phaseB", "value": args.batch * world / (ms.item() / 1000.0), "unit": "images/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms.item(),
                          "config": {"workload": "voc21_dual_student_phaseB_step_448", "per_gpu_batch": args.batch,
                                     "parallelism": f"ddp{world}" if world > 1 else "single"},
                          "loss_first_warmup": losses[0] if losses else None, "loss_last": loss.item(),
                          "parts": {k: float(v) for k, v in parts.items()}, "finite": bool(torch.isfinite(loss).item()),
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step(x, cls, box, 4000)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60), file=sys.stderr)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
