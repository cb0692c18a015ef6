"""CPU, world_size 2 over gloo: the host-side contract between the per-student autograd.Function and
DistributedDataParallel (train_final_voc.py:155: DDP(find_unused_parameters=True)).  The CUDA kernels cannot run
here, so dupl_b200.train._forward/_backward are replaced by deterministic CPU stand-ins; what is tested is the
plumbing that is identical on the GPU: parameters enter the Function as inputs, gradients come back per parameter,
DDP averages them across ranks, `head.*` / `pos_embed` stay unused, both ranks end with identical gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    from dupl_b200 import dense, train
    from dupl_b200.model.model_dupl import siamese_network

    def fake_forward(net, x, size=None, kept=None):
        B = x.shape[0]
        K = net.num_classes - 1
        outs = (torch.zeros(B, K), torch.zeros(B, K + 1, 2, 2), torch.zeros(B, 768, 2, 2), torch.zeros(B, K))
        return outs, object()

    def fake_backward(net, S, g_cls, g_seg, g_x4, g_aux):
        scale = float(g_cls.sum()) if g_cls is not None else 0.0
        return {n: torch.full_like(p, (rank + 1) * scale) for n, p in train.trainable_parameters(net)}

    train._forward, train._backward = fake_forward, fake_backward
    dense._wants_grad = lambda nets: True
    import dupl_b200._lib as L
    L.require_cuda = lambda *t: None
    model = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
    res = ddp(torch.zeros(2, 3, 32, 32))
    assert set(res) == {"branch1", "branch2"}
    loss = res["branch1"][0].sum() * 1.0 + res["branch2"][0].sum() * 2.0 + 0.0 * res["branch1"][1].sum()
    loss.backward()
    g1 = model.branch1.encoder.blocks[3].attn.qkv.weight.grad
    g2 = model.branch2.decoder.conv7.weight.grad
    ok = (model.branch1.encoder.head.weight.grad is None and model.branch1.encoder.pos_embed.grad is None
          and torch.allclose(g1, torch.full_like(g1, 1.5 * 40)) and torch.allclose(g2, torch.full_like(g2, 1.5 * 80)))
    out[rank] = (bool(ok), float(g1.flatten()[0]), float(g2.flatten()[0]))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_student_function_under_ddp_world_size_2():
    port = 29500 + os.getpid() % 1000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out[0][0] and out[1][0], dict(out)
        assert out[0][1:] == out[1][1:]  # identical averaged gradients on both ranks


def _worker_arena(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dupl_b200.grad_arena import GradArena
    names = ["head", "w2", "b2", "w1", "b1", "unused"]
    shapes = [(3, 5), (40, 8), (40,), (64, 4), (64,), (2, 2)]
    params = [(n, torch.nn.Parameter(torch.zeros(s))) for n, s in zip(names, shapes)]
    arena = GradArena(params, names, chunk_elems=300, device=torch.device("cpu"))
    assert len(arena.chunks) >= 2 and arena.flat.numel() % GradArena.ALIGN == 0
    log = []
    orig = arena._reduce

    def spy(lo, hi):
        log.append((lo, hi))
        orig(lo, hi)
    arena._reduce = spy
    for step_no in range(2):
        # phase-C shape of a step: two backward calls reach the arena; the first writes, the second accumulates and releases
        # finished chunks to the (asynchronous) all-reduce as it goes; "unused" never receives a gradient
        arena.begin_step(2)
        for call in range(2):
            for n, s in zip(names[:-1], shapes[:-1]):
                g = torch.full(s, float((rank + 1) * (call + 1) * (step_no + 1)))
                if call == 0 and n == "w1":
                    arena.out(n).copy_(g)            # in-place write into the view, like a wgrad GEMM epilogue
                    g = arena.out(n)
                elif call == 1 and n == "b2":
                    g = None                         # this call contributes nothing to b2
                arena.put(n, g)
            arena.put("unused", None)
            issued_mid = len(log)
            arena.end_call()
        assert issued_mid >= 1                       # chunks were released DURING the last call, before finish()
        arena.finish()
        arena.bind_grads()
    p = dict(params)
    out[rank] = (p["w2"].grad.flatten()[0].item(), p["b2"].grad.flatten()[0].item(), p["w1"].grad.flatten()[0].item(),
                 p["unused"].grad is None, p["w1"].grad.data_ptr() == arena.views["w1"].data_ptr(), log)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_gradient_arena_chunked_all_reduce_averages_over_ranks(world):
    """grad_arena.GradArena (what the captured multi-rank step does instead of DDP's reducer): gradients written / accumulated
    in place, finished chunks all-reduced in arena order while the backward pass is still running, mean over ranks,
    parameters without a gradient keep .grad = None, .grad are views of the arena.  World sizes 2 and 4."""
    port = 29900 + os.getpid() % 90 + world
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_arena, args=(world, port, out), nprocs=world, join=True)
        for r in range(1, world):
            assert out[0][:5] == out[r][:5]
            assert out[0][5] == out[r][5]                 # every rank enqueued the same sequence of collectives
        w2, b2, w1, unused_none, is_view, log = out[0]
        # second step: rank r contributes (r+1)*2*(1 + 2) to w2 / w1 and (r+1)*2*1 to b2; the mean of (r+1) over the ranks
        m = (world + 1) / 2
        assert abs(w2 - m * 6) < 1e-6 and abs(w1 - m * 6) < 1e-6 and abs(b2 - m * 2) < 1e-6
        assert unused_none and is_view
        half = len(log) // 2
        assert log[:half] == log[half:] and [lo for lo, _ in log[:half]] == sorted(lo for lo, _ in log[:half])
