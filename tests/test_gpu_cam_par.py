"""GPU parity tests of the HBM-bound half of the path: MS-CAM post-processing, pseudo-label casts,
PAR and refine — CUDA through the C ABI vs oracle/dupl_oracle.py on identical seeded inputs.
Integer / label outputs: bit-exact.  Float outputs: tolerance stated per test."""
import numpy as np
import pytest
import torch

from helpers import synth_boxes, synth_cls_labels

pytestmark = pytest.mark.gpu

DIL = (1, 2, 4, 8, 12, 24)


def _smooth_img(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (b, 3, h, w), generator=g).float()
    u8 = torch.nn.functional.avg_pool2d(u8, 5, stride=1, padding=2, count_include_pad=False).round()
    return u8 / 255.0


@pytest.mark.parametrize("b,K,H,W,grids", [(2, 20, 64, 64, [(4, 4), (2, 2), (6, 6)]), (1, 5, 96, 60, [(6, 4), (3, 2), (9, 6)]),
                                            (2, 3, 30, 50, [(2, 3)])])
def test_mscam_post(b, K, H, W, grids):
    from dupl_b200 import ops
    from oracle import dupl_oracle as O
    g = torch.Generator().manual_seed(0)
    low = [torch.randn(2 * b, K, gh, gw, generator=g) for gh, gw in grids]
    got = ops.mscam_post([t.cuda() for t in low], b, H, W).cpu()
    want = O.mscam_post(low, b, H, W)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 2e-6
    assert got.min().item() >= 0.0 and got.max().item() < 1.0


@pytest.mark.parametrize("dynamic", [False, True])
@pytest.mark.parametrize("h,w", [(28, 28), (64, 48)])
def test_cam_to_label_bit_exact(dynamic, h, w):
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    b, K = 4, 20
    g = torch.Generator().manual_seed(1)
    cam = torch.rand(b, K, h, w, generator=g)
    cam[0, :, 0, 0] = 0.5  # exact ties -> first index
    cls = synth_cls_labels(b, K, seed=2)
    box = synth_boxes(b, 448, 448, seed=3)  # boxes in 448-pixel coordinates on small maps: slice clipping quirk
    box[1] = torch.tensor([30, 448, 0, 10])  # starts beyond the map -> everything ignored
    ht = torch.tensor([0.6, 0.7, 0.55, 0.65]) if dynamic else 0.65
    fn = cam_helper.cam_to_label_dynamic_cls if dynamic else cam_helper.cam_to_label
    kw = dict(bkg_thre=0.45, low_thre=0.25, ignore_mid=True, ignore_index=255)
    valid, lab = fn(cam.cuda(), cls.cuda(), box, high_thre=ht.cuda() if dynamic else ht, **kw)
    ovalid, olab = O.cam_to_label(cam, cls, box, high_thre=ht, **kw)
    assert lab.dtype == torch.int64 and lab.shape == (b, h, w)
    assert torch.equal(lab.cpu(), olab)
    assert torch.equal(valid.cpu(), ovalid)
    lab2 = cam_helper.cam_to_label(cam.cuda(), cls.cuda(), bkg_thre=0.45)
    assert torch.equal(lab2.cpu(), O.cam_to_label(cam, cls, bkg_thre=0.45))


def test_label_to_aff_mask_bit_exact():
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    g = torch.Generator().manual_seed(4)
    lab = torch.randint(0, 4, (3, 12, 11), generator=g)
    lab[lab == 3] = 255
    got = cam_helper.label_to_aff_mask(lab.cuda())
    assert got.dtype == torch.int64
    assert torch.equal(got.cpu(), O.label_to_aff_mask(lab))


@pytest.mark.parametrize("B,C,h,w", [(1, 3, 40, 56), (2, 5, 64, 33), (1, 1, 25, 25)])
def test_par_forward(B, C, h, w):
    """PAR.forward (PAR.py:64-91): float outputs, tolerance 2e-6 absolute on probabilities."""
    from dupl_b200.model.PAR import PAR
    from oracle import dupl_oracle as O
    img = _smooth_img(B, h, w, seed=5)
    g = torch.Generator().manual_seed(6)
    masks = torch.rand(B, C, h, w, generator=g).softmax(1)
    par = PAR(num_iter=10, dilations=list(DIL)).cuda()
    got = par(img.cuda(), masks.cuda()).cpu()
    want = O.par_forward(img, masks, DIL, 10)
    assert (got - want).abs().max().item() < 2e-6
    aff = par.affinity(img.cuda()).cpu()
    assert (aff - O.par_affinity(img, DIL)).abs().max().item() < 1e-6
    assert torch.allclose(aff.sum(1), torch.full((B, h, w), 1.01), atol=1e-5)  # rows sum to 1 + w2


def test_par_constant_image_has_no_nan():
    from dupl_b200.model.PAR import PAR
    par = PAR(num_iter=2, dilations=[1, 2]).cuda()
    out = par(torch.full((1, 3, 16, 16), 0.5).cuda(), torch.rand(1, 2, 16, 16).cuda())
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("dynamic", [True, False])
def test_refine_labels(dynamic):
    """refine_cams_with_dynamic_thres / _bkg_v2 (cam_helper.py:338-431).
    (1) stage-isolated: epilogue fed the ORACLE's propagated masks is bit-exact;
    (2) end to end: labels may differ only where the arg-max margin is below 1e-5."""
    from dupl_b200 import ops
    from dupl_b200.model.PAR import PAR
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    b, K, H, W = 3, 20, 64, 96
    imgs = _smooth_img(b, H, W, seed=7)
    cls = synth_cls_labels(b, K, seed=8)
    g = torch.Generator().manual_seed(9)
    cams = torch.rand(b, K, H, W, generator=g) * cls[:, :, None, None]
    box = synth_boxes(b, H, W, seed=10)
    bkg = torch.stack([torch.full((1, H, W), v) for v in (0.6, 0.7, 0.55)]) if dynamic else 0.65
    par = PAR(num_iter=10, dilations=list(DIL)).cuda()
    want, want_h, want_l = O.refine_cams(imgs, cams, cls, bkg, 0.25, 255, box, DIL, 10, return_parts=True)
    if dynamic:
        got = cam_helper.refine_cams_with_dynamic_thres(par, imgs.cuda(), cams.cuda(), cls.cuda(), high_thre_map=bkg.cuda(),
                                                        low_thre=0.25, ignore_index=255, img_box=box)
    else:
        got = cam_helper.refine_cams_with_bkg_v2(par, imgs.cuda(), cams.cuda(), cls.cuda(), high_thre=bkg, low_thre=0.25,
                                                 ignore_index=255, img_box=box)
    assert got.dtype == torch.float32 and got.shape == (b, H, W)
    mism = (got.cpu() != want).float().mean().item()
    assert mism < 1e-3, f"label mismatch fraction {mism}"

    # stage-isolated bit-exactness: prologue output vs oracle, and epilogue on oracle masks
    images_ds, masks, nactive, clsd = ops.refine_prologue(imgs.cuda(), cams.cuda(), cls.cuda(),
                                                          bkg.cuda() if dynamic else bkg, 0.25)
    assert torch.equal(images_ds.cpu(), O._down2(imgs))
    present = torch.cat([torch.ones(b, 1, dtype=torch.bool), cls != 0], 1)
    bkg_t = bkg if dynamic else torch.full((b, 1, H, W), float(bkg))
    omasks = torch.zeros(b, 2 * (K + 1), H // 2, W // 2)
    aff = O.par_affinity(O._down2(imgs), DIL)
    for v, bk in enumerate((bkg_t, torch.full((b, 1, H, W), 0.25))):
        mm = O._down2(torch.cat([bk, cams], 1)).masked_fill(~present[:, :, None, None], float("-inf")).softmax(1)
        for i in range(b):
            idx = present[i].nonzero()[:, 0]
            n = len(idx)
            assert int(nactive[i]) == 2 * n
            assert (masks[i, v * n:(v + 1) * n].cpu() - mm[i, idx]).abs().max().item() < 1e-6
            omasks[i, v * n:(v + 1) * n] = O.par_propagate(aff[i:i + 1], mm[i:i + 1, idx], DIL, 10)[0]
    lab, lh, ll = ops.refine_epilogue(omasks.cuda(), clsd, box, H, W, 255, want_parts=True)
    assert torch.equal(lh.cpu(), want_h) and torch.equal(ll.cpu(), want_l) and torch.equal(lab.cpu(), want)


@pytest.mark.parametrize("b,K,H,W,grids", [(2, 20, 448, 448, [(28, 28), (14, 14), (42, 42)]), (1, 5, 64, 96, [(4, 6), (2, 3), (6, 9)]),
                                           (3, 7, 100, 130, [(6, 8)]), (1, 3, 224, 224, [(14, 14), (7, 7)])])
def test_mscam_column_kernel_is_bit_identical_to_the_generic_kernel(b, K, H, W, grids, monkeypatch):
    """The single-launch cluster kernel (values computed once, min / max through distributed shared memory) and the two-pass
    column-per-thread kernels hoist the horizontal interpolation out of the row loop; same fma order => same bits as the
    generic kernel."""
    from dupl_b200 import ops
    g = torch.Generator().manual_seed(5)
    lowres = [torch.randn(2 * b, K, gh, gw, generator=g).cuda() for gh, gw in grids]
    fast = ops.mscam_post(lowres, b, H, W)
    monkeypatch.setenv("DUPL_MSCAM_2PASS", "1")
    two_pass = ops.mscam_post(lowres, b, H, W)
    monkeypatch.setenv("DUPL_MSCAM_GENERIC", "1")
    slow = ops.mscam_post(lowres, b, H, W)
    torch.cuda.synchronize()
    assert torch.isfinite(fast).all()
    assert torch.equal(two_pass, slow)
    assert torch.equal(fast, slow)


def test_mscam_cluster_kernel_short_and_ragged_planes():
    """H smaller than the cluster (CTAs without rows), W not a multiple of 32, a plane that ReLU leaves all-zero."""
    from dupl_b200 import ops
    g = torch.Generator().manual_seed(6)
    for b, K, H, W, grids in [(1, 2, 5, 37, [(2, 3)]), (2, 3, 30, 50, [(3, 5), (2, 2)])]:
        lowres = [torch.randn(2 * b, K, gh, gw, generator=g).cuda() for gh, gw in grids]
        for t in lowres:
            t[0, 0] = -1.0          # image 0, class 0: negative everywhere in the image ...
            t[b, 0] = -2.0          # ... and in its flipped twin -> value 0 everywhere -> 0 / 1e-5
        out = ops.mscam_post(lowres, b, H, W)
        import os
        os.environ["DUPL_MSCAM_GENERIC"] = "1"
        try:
            ref = ops.mscam_post(lowres, b, H, W)
        finally:
            del os.environ["DUPL_MSCAM_GENERIC"]
        assert torch.equal(out, ref)
        assert (out[0, 0] == 0).all()


@pytest.mark.parametrize("B,P,h,w,live", [(4, 42, 224, 224, [4, 6, 8, 10]), (2, 7, 50, 70, None), (1, 3, 24, 33, None),
                                          (3, 12, 64, 64, [12, 1, 5])])
def test_par_tiled_propagation_is_bit_identical_to_the_per_pixel_kernel(B, P, h, w, live, monkeypatch):
    """Shared-memory tile + halo (replicate border baked in) vs direct clamped gathers: same products, same order."""
    from dupl_b200 import ops
    g = torch.Generator().manual_seed(7)
    dil = [1, 2, 4, 8, 12, 24]
    imgs = torch.rand(B, 3, h, w, generator=g).cuda()
    aff = ops.par_affinity(imgs, dil)
    masks = torch.rand(B, P, h, w, generator=g).cuda()
    nact = None if live is None else torch.tensor(live, dtype=torch.int32).cuda()
    fast = ops.par_propagate(aff, masks.clone(), dil, 10, nactive=nact)
    monkeypatch.setenv("DUPL_PAR_SIMPLE", "1")
    slow = ops.par_propagate(aff, masks.clone(), dil, 10, nactive=nact)
    torch.cuda.synchronize()
    for i in range(B):
        n = P if live is None else live[i]
        assert torch.isfinite(fast[i, :n]).all()
        assert torch.equal(fast[i, :n], slow[i, :n])
