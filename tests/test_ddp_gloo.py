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


def _worker_allreduce(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dupl_b200.train_step import TrainStep
    ps = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    ps[0].grad = torch.full((3, 5), float(rank + 1))
    ps[1].grad = torch.arange(7.0) * (rank + 1)
    # ps[2] never received a gradient (like encoder.head.*): it must be skipped, not break the flattening

    class _Opt:
        param_groups = [{"params": ps[:2]}, {"params": ps[2:]}]

    step = TrainStep.__new__(TrainStep)
    step.optim = _Opt()
    orig = dist.all_reduce

    def avg_all_reduce(t, op=None):           # gloo has no AVG: emulate it (NCCL provides it natively)
        orig(t, op=dist.ReduceOp.SUM)
        t.div_(world)
    dist.all_reduce = avg_all_reduce
    step._all_reduce_grads()
    out[rank] = (ps[0].grad.tolist(), ps[1].grad.tolist(), ps[2].grad is None)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_in_graph_gradient_all_reduce_averages_over_ranks():
    """TrainStep._all_reduce_grads (what the captured multi-rank step does instead of DDP's reducer): mean over ranks,
    parameters without a gradient are skipped."""
    port = 29900 + os.getpid() % 90
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_allreduce, args=(2, port, out), nprocs=2, join=True)
        assert out[0] == out[1]
        g0, g1, none2 = out[0]
        assert none2 and all(abs(v - 1.5) < 1e-6 for row in g0 for v in row)
        assert all(abs(v - 1.5 * i) < 1e-6 for i, v in enumerate(g1))
