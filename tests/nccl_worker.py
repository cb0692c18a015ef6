"""Worker of tests/test_gpu_ddp_nccl.py (one process per GPU under torchrun, backend nccl).

Checks on real NCCL what the gloo tests check on stand-ins:
  1. the arena path (gradients written in place + chunked all-reduce overlapped with the backward pass) leaves in every
     .grad the MEAN over the ranks of the per-rank gradients — compared with gradients this rank computes eagerly for both
     ranks' batches (train_final_voc.py:155,470: DistributedDataParallel's semantics);
  2. after k captured steps (whole iteration as one CUDA graph, NCCL inside) ALL parameters are bit-identical on all ranks
     and moved away from their initial values.
Prints one JSON line on rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images  # noqa: E402


def batch(seed, b, S):
    return synth_images(b, S, S, seed=seed).cuda(), synth_cls_labels(b, 20, seed=seed).cuda(), synth_boxes(b, S, S, seed=seed)


def mark(msg):
    if os.environ.get("DUPL_TEST_TRACE"):
        print(f"[rank {os.environ.get('RANK')}] {msg}", file=sys.stderr, flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DUPL_TEST_WATCHDOG", "240")), exit=True)   # a hang names its line
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dupl_b200.model.model_dupl import siamese_network
    from dupl_b200.train_step import TrainStep, make_optimizer
    S, b, n_iter = int(os.environ.get("DUPL_TEST_SIZE", "64")), 2, 3000
    P = init_state_dict(21)

    def model():
        m = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
        m.load_state_dict(P, strict=True)
        return m.cuda().train()

    # ---- 1. expected mean gradient, computed eagerly on this rank for every rank's batch
    m = model()
    eager = TrainStep(m, None, graph=False)
    expect = None
    for r in range(world):
        m.zero_grad(set_to_none=True)
        loss, _, _ = eager.losses(*batch(100 + r, b, S), n_iter)
        loss.backward()
        g = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
        expect = g if expect is None else {n: expect[n] + g[n] for n in g}
    expect = {n: v / world for n, v in expect.items()}
    mark("expected gradients done")
    # ---- arena path on this rank's own batch (chunk size small enough for several chunks even at test size)
    os.environ.setdefault("DUPL_GRAD_CHUNK_ELEMS", str(4 << 20))
    os.environ.setdefault("DUPL_GRAD_OVERLAP", "1")     # the chunked / overlapped path with its own NCCL group (default from 4 ranks)
    m2 = model()
    opt = make_optimizer(m2, capturable=True)
    step = TrainStep(m2, opt, capture=True)
    x, cls, box = batch(100 + rank, b, S)
    mark("arena step constructed")
    loss, _, _ = step.losses(x, cls, box, n_iter)
    mark("arena losses done")
    step._backward(loss)
    mark("arena backward issued")
    torch.cuda.synchronize()
    mark("arena backward complete")
    worst, missing = 0.0, []
    for n, p in m2.named_parameters():
        if n in expect:
            if p.grad is None:
                missing.append(n)
                continue
            e = ((p.grad.double() - expect[n].double()).norm() / expect[n].double().norm().clamp_min(1e-30)).item()
            worst = max(worst, e)
        elif p.grad is not None:
            missing.append("unexpected:" + n)
    chunks = [len(a.chunks) for a in step._arenas]
    del loss
    # ---- 2. k captured steps, then every parameter bit-identical across ranks
    init = {n: p.detach().clone() for n, p in m2.named_parameters()}
    losses = []
    for i in range(3):
        l, _ = step(x, cls, box, n_iter + i)
        losses.append(float(l))
        mark(f"captured step {i} done")
    torch.cuda.synchronize()
    chk = torch.stack([p.detach().contiguous().view(torch.int32).to(torch.int64).sum() for p in m2.parameters()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    moved = sum(1 for n, p in m2.named_parameters() if p.requires_grad and ".head." not in n and not torch.equal(p, init[n]))
    w = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("RESULT " + json.dumps({"world": world, "grad_mean_worst_rel": float(w), "missing": missing[:5], "chunks": chunks,
                                      "params_checked": int(chk.numel()), "params_in_sync": bool((lo == hi).all()),
                                      "params_moved": moved, "losses": losses}))
    # the captured graphs hold NCCL work of the gradient group: release them before the communicators go away
    # (destroy_process_group() blocks forever on a communicator that a live CUDA graph still references)
    import gc
    del step, opt, m2, m, eager
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
