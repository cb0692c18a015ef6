"""GPU: DenseCRF mean-field (dupl_b200.utils.dcrf, crf.cu) vs the C restatement oracle/densecrf_ref.c on
identical inputs.  PARITY UNPINNED w.r.t. pydensecrf itself (third-party, absent) — see DESIGN.md §5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(H, W, C, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1).float()[None]
    img = torch.nn.functional.avg_pool2d(t, 7, 1, 3, count_include_pad=False)[0].permute(1, 2, 0).round().numpy().astype(np.uint8)
    lg = rng.randn(C, H // 8 + 1, W // 8 + 1).astype(np.float32) * 2
    lg = torch.nn.functional.interpolate(torch.from_numpy(lg)[None], size=(H, W), mode="bilinear", align_corners=False)[0].numpy()
    p = np.exp(lg - lg.max(0, keepdims=True))
    p /= p.sum(0, keepdims=True)
    return np.ascontiguousarray(img), p.astype(np.float32)


@pytest.mark.parametrize("H,W,C,params", [
    (48, 64, 5, (10, 1, 1, 4, 121, 5)),      # eval_seg_voc.py:104-111 parameters
    (60, 45, 21, (10, 3, 3, 10, 80, 13)),    # crf_inference parameters (dcrf.py:18-19)
    (33, 47, 81, (5, 1, 1, 4, 121, 5)),      # COCO class count (3 classes per lane)
    (40, 40, 4, (3, 3, 3, 0, 10, 10)),       # Gaussian only
    (40, 40, 33, (3, 0, 3, 5, 20, 8)),       # bilateral only
])
def test_crf_matches_c_restatement(H, W, C, params):
    from dupl_b200.utils.dcrf import DenseCRF
    from oracle.densecrf_ref import DenseCRF as RefCRF
    img, p = _case(H, W, C, seed=H * W + C)
    ref = RefCRF(*params)
    want = ref(img, p)
    got = DenseCRF(*params)(img, p)
    assert got.dtype == np.float32 and got.shape == (C, H, W)
    assert np.abs(got.sum(0) - 1).max() < 1e-5
    assert np.abs(got - want).max() < 2e-5
    agree = (got.argmax(0) == want.argmax(0)).mean()
    assert agree > 0.9995
    # the lattices themselves are identical (vertex counts)
    from dupl_b200 import ops
    _, sizes = ops.crf_inference(torch.from_numpy(img).cuda(), torch.from_numpy(p).cuda(), params[0], *[float(x) for x in params[1:]])
    for k, w in enumerate((params[1], params[3])):
        if w != 0:
            assert sizes[k] == int(ref.lattice_sizes[k])


def test_crf_is_bit_reproducible():
    from dupl_b200.utils.dcrf import DenseCRF
    img, p = _case(64, 80, 21, seed=3)
    crf = DenseCRF(10, 1, 1, 4, 121, 5)
    a = crf(img, p)
    b = crf(img, p)
    assert np.array_equal(a, b)


def test_crf_label_variants_run_and_agree_with_oracle():
    from dupl_b200.utils import dcrf
    from oracle import densecrf_ref as R
    img, p = _case(40, 56, 21, seed=4)
    q = dcrf.crf_inference(img, p, t=5, labels=21)
    want = R.DenseCRF(5, 3, 3, 10, 80, 13)(img, p)
    assert np.abs(q - want).max() < 2e-5
    lab = p.argmax(0)
    out = dcrf.crf_inference_label(img, lab, t=3, n_labels=21, gt_prob=0.7)
    assert out.shape == lab.shape and out.dtype == np.int64
    assert (out == lab).mean() > 0.5


def test_numpy_interface_through_the_pinned_staging_buffers():
    """A probability map above 256 K elements takes the page-locked staging path (utils/dcrf.py): numpy in / numpy out must
    equal the device-tensor path bit for bit, and every call must hand back its own array (the staging buffers are reused)."""
    from dupl_b200.utils.dcrf import DenseCRF
    img, p = _case(160, 200, 21, seed=5)
    img2, p2 = _case(160, 200, 21, seed=6)
    assert p.size > (1 << 18)
    crf = DenseCRF(5, 1, 1, 4, 121, 5)
    a = crf(img, p)
    keep = a.copy()
    b = crf(img2, p2)
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == p.shape
    assert np.array_equal(a, keep) and not np.array_equal(a, b)        # the second call did not overwrite the first result
    dev = crf(torch.from_numpy(img).cuda(), torch.from_numpy(p).cuda())
    assert torch.is_tensor(dev) and dev.is_cuda
    assert np.array_equal(dev.cpu().numpy(), a)
