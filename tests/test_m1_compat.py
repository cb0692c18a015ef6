"""Driver M1 plumbing that can be checked without a GPU: the compat layer imports the unmodified reference from
baseline/_ref (or /root/reference), the drop-in modules bind under the names the scripts import, and the drop-in
`siamese_network(pretrained=True)` starts from bit-identical parameters as the reference's under the same seed
(train_final_voc.py:95-102,145-150) — the precondition for comparing the two script runs iteration by iteration."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import compat  # noqa: E402

if not compat.available() and os.path.isdir("/root/reference/model"):
    compat.REF_ROOT = "/root/reference"

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not compat.available(), reason="no reference tree (baseline/_ref)")]


def test_same_seed_gives_bit_identical_initial_parameters_with_offline_pretrained_weights():
    from dupl_b200.model.model_dupl import siamese_network as ours
    ref = compat.load_reference()
    compat.patch_hub_offline()
    torch.manual_seed(0)
    a = ref.model_dupl.siamese_network(backbone="deit_base_patch16_224", num_classes=21, pretrained=True, aux_layer=-3)
    torch.manual_seed(0)
    b = ours(backbone="deit_base_patch16_224", num_classes=21, pretrained=True, aux_layer=-3)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    # and without pretrained weights: the constructors draw the same random numbers in the same order
    torch.manual_seed(3)
    a = ref.model_dupl.siamese_network(backbone="deit_base_patch16_224", num_classes=81, pretrained=False, aux_layer=9)
    torch.manual_seed(3)
    b = ours(backbone="deit_base_patch16_224", num_classes=81, pretrained=False, aux_layer=9)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert [len(g) for g in a.get_param_groups()] == [len(g) for g in b.get_param_groups()]
    for ga, gb in zip(a.get_param_groups(), b.get_param_groups()):
        for pa, pb in zip(ga, gb):
            assert pa.shape == pb.shape


def test_scripts_import_the_dropin_modules_under_the_reference_names(tmp_path):
    """baseline/run_script.py --dropin: the script's import statements (train_final_voc.py:19-24) resolve to dupl_b200;
    datasets / optimizer / logging stay the reference's.  Run in a subprocess (the binding edits sys.modules)."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from baseline import compat\n"
        "compat.REF_ROOT = %r\n"
        "compat.enter_reference_tree(dropin=True)\n"
        "from datasets import voc\n"
        "from model.losses import get_masked_ptc_loss, get_seg_loss\n"
        "from model.model_dupl import siamese_network\n"
        "from model.PAR import PAR\n"
        "from utils import evaluate, imutils, cam_helper, train_helper, pyutils\n"
        "from utils.dcrf import DenseCRF\n"
        "mods = [siamese_network.__module__, PAR.__module__, get_seg_loss.__module__, cam_helper.__name__, DenseCRF.__module__,\n"
        "        train_helper.cam_helper.__name__, voc.__file__, train_helper.__file__]\n"
        "print('|'.join(mods))\n" % (ROOT, compat.REF_ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    mods = out.stdout.strip().splitlines()[-1].split("|")
    assert mods[:6] == ["dupl_b200.model.model_dupl", "dupl_b200.model.PAR", "dupl_b200.model.losses", "dupl_b200.utils.cam_helper",
                        "dupl_b200.utils.dcrf", "dupl_b200.utils.cam_helper"]
    assert mods[6].startswith(compat.REF_ROOT) and mods[7].startswith(compat.REF_ROOT)


def test_synthetic_dataset_feeds_the_reference_loader(tmp_path):
    """baseline/make_dataset.py writes what datasets/voc.py:36-39,95 reads; one batch through the reference's own
    VOC12ClsDataset + DataLoader has the shapes / dtypes of train_final_voc.py:181."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from baseline import compat, make_dataset\n"
        "compat.REF_ROOT = %r\n"
        "root, lists = make_dataset.make_voc_like(%r, n_train=4, n_val=1)\n"
        "compat.enter_reference_tree(dropin=False)\n"
        "from datasets import voc\n"
        "from torch.utils.data import DataLoader\n"
        "ds = voc.VOC12ClsDataset(root_dir=root, name_list_dir=lists, split='train_aug', stage='train', aug=True, rescale_range=(0.5, 2),\n"
        "                         crop_size=448, img_fliplr=True, ignore_index=255, num_classes=21)\n"
        "name, x, cls, box, crops = next(iter(DataLoader(ds, batch_size=2, num_workers=0, drop_last=True)))\n"
        "print(tuple(x.shape), x.dtype, tuple(cls.shape), cls.dtype, tuple(box.shape), box.dtype)\n" % (ROOT, compat.REF_ROOT, str(tmp_path)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == "(2, 3, 448, 448) torch.float32 (2, 20) torch.float32 (2, 4) torch.int16"
