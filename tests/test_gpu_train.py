"""GPU: training-mode forward + backward of the students (dupl_b200.train, autograd.Function over CUDA kernels)
against the oracle's CPU autograd: every parameter gradient within 1e-3 (norm-relative) of fp32."""
import pytest
import torch

from helpers import init_state_dict, rel_err, synth_images

pytestmark = pytest.mark.gpu


def _models(num_classes=21):
    from dupl_b200.model.model_dupl import siamese_network
    P = init_state_dict(num_classes)
    m = siamese_network("deit_base_patch16_224", num_classes=num_classes, pretrained=False, aux_layer=-3)
    m.load_state_dict(P, strict=True)
    return m.cuda().train(), P


def _probe_weights(outs, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(o.shape, generator=g) for o in outs]


def _nrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("H,W", [(32, 48), (64, 64)])
def test_student_forward_backward_matches_oracle_autograd(H, W):
    from oracle import dupl_oracle as O
    m, P = _models()
    x = synth_images(2, H, W, seed=7)
    # ---- oracle (CPU autograd) for student 2
    Pg = {k: v.clone().requires_grad_(k.startswith("branch2.") and "pos_embed" not in k and ".head." not in k) for k, v in P.items()}
    want = O.network_forward(Pg, 2, x)
    probes = _probe_weights(want, seed=H)
    sum(( o * w).sum() for o, w in zip(want, probes)).backward()
    # ---- CUDA
    got = m(x.cuda(), branch=2)
    assert len(got) == 4
    for g, w in zip(got, want):
        assert g.shape == w.shape and rel_err(g, w) < 1e-3
    sum((o * w.cuda()).sum() for o, w in zip(got, probes)).backward()
    checked = 0
    worst = (0.0, "")
    bad = []
    for name, p in m.branch2.named_parameters():
        ref = Pg["branch2." + name].grad
        if name.startswith("encoder.head.") or name == "encoder.pos_embed":
            assert p.grad is None
            continue
        assert p.grad is not None, name
        assert torch.isfinite(p.grad).all(), name
        e = _nrel(p.grad, ref)
        worst = max(worst, (e, name))
        if e >= 1e-3:
            bad.append((name, round(e, 5)))
        checked += 1
    assert checked == 154
    assert not bad, bad
    assert all(p.grad is None for p in m.branch1.parameters())  # the other student was not touched


def test_in_place_operands_give_the_gradients_of_the_transposed_copies(monkeypatch):
    """DUPL_MN_MAJOR=1 (default: wgrad / dgrad GEMMs read dY, the saved activations and the weights in place as MN-major
    tcgen05 operands) vs =0 (transposed copies): the same products; only the split-K grouping of a few wgrad shapes may
    differ (fp32 regrouping), everything without split-K is bit-identical."""
    m, _ = _models()
    x = synth_images(2, 64, 64, seed=21).cuda()
    grads = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("DUPL_MN_MAJOR", mode)
        m.zero_grad(set_to_none=True)
        outs = m(x)
        probes = _probe_weights(outs["branch1"], seed=5)
        sum((o * w.cuda()).sum() for br in ("branch1", "branch2") for o, w in zip(outs[br], probes)).backward()
        grads[mode] = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    assert grads["0"].keys() == grads["1"].keys() and len(grads["1"]) == 308
    worst = max((_nrel(grads["1"][n], grads["0"][n]), n) for n in grads["0"])
    assert worst[0] < 2e-6, worst
    same = sum(torch.equal(grads["1"][n], grads["0"][n]) for n in grads["0"])
    assert same >= 100, same          # biases / norms / classifier heads never see a GEMM regrouping


def test_both_students_dict_output_and_need_sp_view():
    """model(x) -> {'branch1': 4-tuple, 'branch2': 4-tuple}; need_sp adds the 0.75x aug-view seg logits (model_dupl.py:190-205)."""
    from oracle import dupl_oracle as O
    m, P = _models()
    x = synth_images(2, 64, 64, seed=8)
    x_aug = synth_images(2, 64, 64, seed=9)
    res = m(torch.cat([x, x_aug]).cuda(), need_sp=True)
    assert set(res) == {"branch1", "branch2", "branch1_aug", "branch2_aug"}
    with torch.no_grad():
        want1 = O.network_forward(P, 1, x)
        small = O.bilinear(x_aug, 48, 48)
        want_aug = O.network_forward(P, 2, small)[1]
    for g, w in zip(res["branch1"], want1):
        assert rel_err(g, w) < 1e-3
    assert res["branch2_aug"].shape == want_aug.shape == (2, 21, 3, 3)
    assert rel_err(res["branch2_aug"], want_aug) < 1e-3
    (res["branch1"][1].sum() + res["branch2_aug"].sum()).backward()
    assert m.branch1.decoder.conv8.weight.grad is not None and m.branch2.decoder.conv8.weight.grad is not None


def test_phase_b_step_losses_and_gradients_match_oracle_loop():
    """dupl_b200.train_step.PhaseBStep (restatement of train_final_voc.py:260-356,438-456) vs the oracle's CPU loop:
    every loss part within 1e-3, refined labels (near-)identical, gradient of the total loss within 1e-3."""
    from dupl_b200.train_step import PhaseBStep, VOC_HIGH_THRES_TARGET
    from helpers import synth_boxes, synth_cls_labels
    from oracle import dupl_oracle as O
    m, P = _models()
    b, S = 2, 64
    x = synth_images(b, S, S, seed=11)
    cls = synth_cls_labels(b, 20, seed=12)
    box = synth_boxes(b, S, S, seed=13)
    Pg = {k: v.clone().requires_grad_("pos_embed" not in k and ".head." not in k) for k, v in P.items()}
    want, wparts, wlabels = O.phase_b_losses(Pg, x, cls, box, 3000, thres_target=VOC_HIGH_THRES_TARGET)
    want.backward()
    step = PhaseBStep(m, None)
    got, parts, labels = step.losses(x.cuda(), cls.cuda(), box, 3000)
    got.backward()
    for k in wparts:
        assert abs(parts[k].item() - wparts[k].item()) < 1e-3 * max(1.0, abs(wparts[k].item())), (k, parts[k].item(), wparts[k].item())
    for a, w in zip(labels, wlabels):
        assert (a.cpu() != w).float().mean().item() < 1e-3
    bad = []
    for name, p in m.named_parameters():
        ref = Pg[name].grad
        if ".head." in name or "pos_embed" in name:
            continue
        e = _nrel(p.grad, ref)
        if e >= 2e-3:
            bad.append((name, round(e, 5)))
    assert not bad, bad[:10]


def test_forward_reuse_from_the_mscam_pass_is_bit_identical():
    """PhaseBStep(reuse_forward=True) starts the training pass from the activations the MS-CAM pass kept for the
    un-flipped scale-1.0 images instead of running model(inputs) again: losses and every gradient must be bit-equal."""
    from dupl_b200.train_step import PhaseBStep
    from helpers import synth_boxes, synth_cls_labels
    b, S = 2, 64
    x = synth_images(b, S, S, seed=21).cuda()
    cls = synth_cls_labels(b, 20, seed=22).cuda()
    box = synth_boxes(b, S, S, seed=23)
    results = []
    for reuse, graph in ((False, False), (True, False), (True, True)):
        m, _ = _models()
        step = PhaseBStep(m, None, graph=graph, reuse_forward=reuse)
        for _ in range(2):  # second call replays the captured graph
            m.zero_grad(set_to_none=True)
            loss, parts, labels = step.losses(x, cls, box, 3000)
            loss.backward()
        torch.cuda.synchronize()
        results.append((loss.detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
    base_loss, base_grads = results[0]
    for loss, grads in results[1:]:
        assert torch.equal(loss, base_loss)
        assert grads.keys() == base_grads.keys()
        for n in grads:
            assert torch.equal(grads[n], base_grads[n]), n


def _check_step(m, P, cfg, args, n_iter, num_classes, cls_dtype=torch.float32, with_aug=False, seed=50):
    from dupl_b200.train_step import TrainStep
    from helpers import synth_boxes, synth_cls_labels, synth_images
    from oracle import dupl_oracle as O
    b, S = 2, 64
    x = synth_images(b, S, S, seed=seed)
    x_aug = synth_images(b, S, S, seed=seed + 1) if with_aug else None
    cls = synth_cls_labels(b, num_classes - 1, seed=seed + 2).to(cls_dtype)
    box = synth_boxes(b, S, S, seed=seed + 3)
    Pg = {k: v.clone().requires_grad_("pos_embed" not in k and ".head." not in k) for k, v in P.items()}
    want, wparts, wlabels = O.train_losses(Pg, x, cls, box, n_iter, cfg, thres_target=list(args.high_thres_target), inputs_aug=x_aug)
    want.backward()
    step = TrainStep(m, None, args=args)
    got, parts, labels = step.losses(x.cuda(), cls.cuda(), box, n_iter, inputs_aug=None if x_aug is None else x_aug.cuda())
    got.backward()
    for k in wparts:
        assert abs(float(parts[k]) - float(wparts[k])) < 1e-3 * max(1.0, abs(float(wparts[k]))), (k, float(parts[k]), float(wparts[k]))
    assert abs(got.item() - want.item()) < 1e-3 * max(1.0, abs(want.item()))
    if wlabels is None:
        assert labels is None
    else:
        for a, w in zip(labels, wlabels):
            assert (a.cpu() != w).float().mean().item() < 1e-3
    bad = []
    for name, p in m.named_parameters():
        ref = Pg[name].grad
        if ".head." in name or "pos_embed" in name:
            continue
        if ref is None or ref.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max() == 0, name
            continue
        e = _nrel(p.grad, ref)
        if e >= 2e-3:
            bad.append((name, round(e, 5)))
    assert not bad, bad[:10]


def test_phase_a_step_matches_oracle_loop():
    """n_iter < cam_iters (train_final_voc.py:194-258): CAM + cls + PTC (static thresholds) + discrepancy loss, no refine."""
    from dupl_b200.train_step import Args
    from oracle import dupl_oracle as O
    m, P = _models()
    _check_step(m, P, O.VOC_CFG, Args, 500, 21)


def test_phase_c_step_matches_oracle_loop():
    """n_iter >= gmm_iters (train_final_voc.py:290-436): need_sp forward of the augmented view, GMM noise filter (GPU kernel vs
    sklearn in the oracle), consistency term."""
    from dupl_b200.train_step import Args
    from oracle import dupl_oracle as O
    m, P = _models()
    _check_step(m, P, O.VOC_CFG, Args, 9000, 21, with_aug=True)


@pytest.mark.parametrize("n_iter", [10000, 20000])
def test_coco_step_matches_oracle_loop(n_iter):
    """train_final_coco.py: 81 classes, uint8 cls_label, aux_layer=9; n_iter <= 12000 refines cams_aux with the scalar
    threshold (refine_cams_with_bkg_v2, :312-322), later the dynamic variant; COCO loss weights (:441-448)."""
    from dupl_b200.model.model_dupl import siamese_network
    from dupl_b200.train_step import CocoArgs
    from oracle import dupl_oracle as O
    P = init_state_dict(81)
    m = siamese_network("deit_base_patch16_224", num_classes=81, pretrained=False, aux_layer=9)
    m.load_state_dict(P, strict=True)
    m = m.cuda().train()
    _check_step(m, P, O.COCO_CFG, CocoArgs, n_iter, 81, cls_dtype=torch.uint8, seed=60)


@pytest.mark.parametrize("fused", [True, False])
def test_captured_iteration_matches_the_eager_iteration(fused):
    """TrainStep(capture=True): the whole iteration as one CUDA graph (gradients in the flat arenas, both students' backward in
    lock step, fused multi-tensor AdamW or torch's capturable AdamW).  Three steps (capture + 2 replays) must leave the same
    parameters as three eager steps with torch.optim.AdamW: same forward / backward kernels; the optimizers differ by fp32
    rounding only (torch's capturable AdamW evaluates the bias corrections in fp32 on the device: 1 - 0.999f is 4.7e-5 off
    1e-3, i.e. ~2e-5 relative on the size of the first updates)."""
    from dupl_b200.train_step import TrainStep, make_optimizer
    from helpers import synth_boxes, synth_cls_labels
    b, S = 2, 64
    x = synth_images(b, S, S, seed=71).cuda()
    cls = synth_cls_labels(b, 20, seed=72).cuda()
    box = synth_boxes(b, S, S, seed=73)
    finals, losses = [], []
    for capture in (False, True):
        m, _ = _models()
        optim = make_optimizer(m, capturable=capture, fused=fused and capture)
        step = TrainStep(m, optim, capture=capture)
        ls = []
        for i in range(3):
            loss, parts = step(x, cls, box, 3000 + i)
            ls.append(loss.item())
        torch.cuda.synchronize()
        losses.append(ls)
        finals.append({n: p.detach().clone() for n, p in m.named_parameters()})
    for a, b_ in zip(*losses):
        assert abs(a - b_) < 1e-4 * max(1.0, abs(a)), losses
    assert abs(losses[0][0] - losses[1][0]) < 1e-6 and abs(losses[0][1] - losses[1][1]) < 2e-6   # before any sizeable update: identical
    assert losses[1][2] != losses[1][0]                      # the replays really train (the forward sees the updated weights)
    worst = max(_nrel(finals[1][n], finals[0][n]) for n in finals[0])
    assert worst < 1e-4, worst


def test_alternating_phases_with_graph_replay_use_their_own_kept_activations():
    """TrainStep owns three graph-replayed CamParStep variants on ONE model; each replay must hand the training forward
    the activations IT wrote.  Phase B -> A -> B with other images each time (graph=True, forward reuse on) must give the
    same losses and gradients as a fresh eager step without reuse on the same inputs."""
    from dupl_b200.train_step import TrainStep
    from helpers import synth_boxes, synth_cls_labels
    b, S = 2, 64
    m, _ = _models()
    step = TrainStep(m, None, graph=True, reuse_forward=True)
    plain_m, _ = _models()
    plain = TrainStep(plain_m, None, graph=False, reuse_forward=False)
    seq = [(3000, 81), (500, 82), (3001, 83), (501, 84), (3002, 85)]
    for n_iter, seed in seq:
        x = synth_images(b, S, S, seed=seed).cuda()
        cls = synth_cls_labels(b, 20, seed=seed).cuda()
        box = synth_boxes(b, S, S, seed=seed)
        outs = []
        for st, mod in ((step, m), (plain, plain_m)):
            mod.zero_grad(set_to_none=True)
            loss, parts, _ = st.losses(x, cls, box, n_iter)
            loss.backward()
            torch.cuda.synchronize()
            outs.append((loss.detach().clone(), {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}))
        assert torch.equal(outs[0][0], outs[1][0]), (n_iter, outs[0][0].item(), outs[1][0].item())
        for n in outs[1][1]:
            assert torch.equal(outs[0][1][n], outs[1][1][n]), (n_iter, n)


def test_phase_c_augments_on_the_device_and_runs_captured():
    """n_iter >= gmm_iters without an explicit augmented view: TrainStep draws the RandAugment operations like the script
    (train_final_voc.py:190-191) and augments on the GPU; same losses as handing the same view over explicitly; the captured
    iteration refills the operation indices per step."""
    import random
    from dupl_b200.pipeline import denormalize_img2
    from dupl_b200.train_step import TrainStep, make_optimizer
    from dupl_b200.utils import imutils
    from helpers import synth_boxes, synth_cls_labels
    b, S = 2, 64
    x = synth_images(b, S, S, seed=91).cuda()
    cls = synth_cls_labels(b, 20, seed=92).cuda()
    box = synth_boxes(b, S, S, seed=93)
    m, _ = _models()
    step = TrainStep(m, None)
    random.seed(3)
    own, _, _ = step.losses(x, cls, box, 9000)
    random.seed(3)
    aug = imutils.augment_data_strong(denormalize_img2(x.clone()), n=5, m=10)
    explicit, _, _ = step.losses(x, cls, box, 9000, inputs_aug=aug)
    assert torch.equal(own.detach(), explicit.detach())
    assert not torch.equal(aug, x)
    m2, _ = _models()
    cap = TrainStep(m2, make_optimizer(m2, capturable=True), capture=True)
    ls = [cap(x, cls, box, 9000 + i)[0].item() for i in range(3)]
    assert all(torch.isfinite(torch.tensor(ls))) and len(set(ls)) > 1      # different operations / updated weights per step


@pytest.mark.parametrize("R,Cc,gelu", [(3140, 768, False), (3140, 3072, True), (3140, 2304, False), (3136, 512, False),
                                        (34, 768, False), (130, 3072, True)])
def test_vectorised_split_kernel_matches_the_tile_kernel(R, Cc, gelu):
    """split_rows_kernel (no transposed output: float4 loads, column sums finished by the last CTA of a strip) vs
    split_transpose_kernel (64x32 tiles, also writes the transposed planes): identical planes and identical bias sums —
    same row -> warp assignment and the same order of the partial sums — and the sums agree with fp64."""
    from dupl_b200 import train
    g = torch.Generator().manual_seed(R + Cc)
    src = torch.randn(R, Cc, generator=g).cuda()
    pre = torch.randn(R, Cc, generator=g).cuda() if gelu else None
    (hi, lo), _, cs = train.split_transpose(src, R, Cc, want_t=False, want_colsum=True, gelu_pre=pre)
    (hi2, lo2), (thi, tlo), cs2 = train.split_transpose(src, R, Cc, want_t=True, want_colsum=True, gelu_pre=pre)
    torch.cuda.synchronize()
    assert torch.equal(hi, hi2) and torch.equal(lo, lo2)
    assert torch.equal(cs, cs2)
    assert torch.equal(thi[:, :R].t().contiguous(), hi) and torch.equal(tlo[:, :R].t().contiguous(), lo)
    ref = src.double()
    if gelu:
        p = pre.double()
        ref = ref * (0.5 * (1 + torch.erf(p * 0.7071067811865476)) + p * torch.exp(-0.5 * p * p) * 0.3989422804014327)
    assert float((cs.double() - ref.sum(0)).abs().max() / ref.sum(0).abs().max()) < 2e-6
    assert rel_err(hi.float() + lo.float(), ref) < 1e-4 if gelu else rel_err(hi.float() + lo.float(), ref) < 2e-5
