"""CPU: the files that travel between the reference's tools and this package: the DDP-prefixed
checkpoint `train_final_voc.py:508` saves loads strict=True into the drop-in model; the `{"msc_seg": ...}` .npy hand-off and the
label PNGs are byte-compatible with what tools/eval_seg_voc.py:83-84,116-117,140 write and read."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import compat  # noqa: E402

if not compat.available() and os.path.isdir("/root/reference/model"):
    compat.REF_ROOT = "/root/reference"


@pytest.mark.reference
@pytest.mark.skipif(not compat.available(), reason="no reference tree")
def test_ddp_checkpoint_of_the_reference_loads_strict_into_the_dropin_model(tmp_path):
    from dupl_b200.eval_sweep import load_checkpoint
    from dupl_b200.model.model_dupl import siamese_network
    ref = compat.load_reference()
    torch.manual_seed(5)
    theirs = ref.model_dupl.siamese_network(backbone="deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    # what `torch.save(model.state_dict(), ckpt_name)` writes for the DistributedDataParallel wrapper (train_final_voc.py:155,508)
    ckpt = {"module." + k: v for k, v in theirs.state_dict().items()}
    path = tmp_path / "checkpoint.pth"
    torch.save(ckpt, path)
    ours = siamese_network(backbone="deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    load_checkpoint(ours, str(path))
    for k, v in theirs.state_dict().items():
        assert torch.equal(ours.state_dict()[k], v), k
    # a missing or extra key must fail (strict=True), as in the tools
    bad = dict(ckpt)
    bad.pop("module.branch1.decoder.conv8.weight")
    with pytest.raises(RuntimeError):
        load_checkpoint(ours, bad)


def test_msc_seg_npy_and_label_png_formats(tmp_path):
    from PIL import Image
    from dupl_b200.eval_sweep import load_msc_seg, save_label_png, save_msc_seg
    g = torch.Generator().manual_seed(0)
    seg = torch.randn(1, 21, 30, 40, generator=g)
    p = str(tmp_path / "2007_000033.npy")
    save_msc_seg(p, seg)
    # read back exactly as crf_proc._job does (tools/eval_seg_voc.py:116-117)
    theirs = np.load(p, allow_pickle=True).item()["msc_seg"]
    assert theirs.dtype == np.float32 and theirs.shape == (1, 21, 30, 40) and np.array_equal(theirs, seg.numpy())
    # a file written the tool's way (tools/eval_seg_voc.py:83) reads back through load_msc_seg
    p2 = str(tmp_path / "tool.npy")
    np.save(p2, {"msc_seg": seg.numpy()})
    assert torch.equal(load_msc_seg(p2), seg)
    pred = torch.randint(0, 21, (30, 40))
    png = str(tmp_path / "pred.png")
    save_label_png(png, pred)
    back = np.asarray(Image.open(png))
    assert back.dtype == np.uint8 and np.array_equal(back, pred.numpy().astype(np.uint8))
