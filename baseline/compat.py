"""What the UNMODIFIED reference tree (baseline/_ref, a verbatim copy of Wu0409/DuPL made by baseline/install_ref.sh)
needs from its environment to run in this image, offline.

Nothing here re-implements reference arithmetic.  It provides
  * in-memory stand-ins for third-party modules the reference imports but this image lacks (timm: 5 symbols touched at
    import time; matplotlib.pyplot; tensorboardX.SummaryWriter; imageio.imread/imsave over PIL; texttable.Texttable);
  * an offline `torch.hub.load_state_dict_from_url` (the scripts default to --pretrained True and download DeiT-B,
    model/backbone/deit.py:102-108): a seeded state dict with the DeiT-B/16 schema, identical for every caller;
  * `load_reference()` — imports the reference's own modules from baseline/_ref without letting its top-level package
    names `model` / `utils` / `datasets` leak into sys.modules;
  * `enter_reference_tree(dropin=...)` — for running the reference SCRIPTS: puts baseline/_ref first on sys.path and,
    with dropin=True, binds the module names the scripts import (`model.model_dupl`, `model.PAR`, `model.losses`,
    `utils.cam_helper`, `utils.camutils`, `utils.dcrf`) to dupl_b200's drop-in modules — what the one-line stubs of
    INTEGRATION.md do inside a maintainer's tree.

Used by bench.py's reference arms, baseline/run_script.py (driver M1) and the tests; never by dupl_b200/.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DUPL_REFERENCE_ROOT", os.path.join(HERE, "_ref"))
REF_PACKAGES = ("model", "utils", "datasets", "tools")
DROPIN_MODULES = ("model.model_dupl", "model.PAR", "model.losses", "utils.cam_helper", "utils.camutils", "utils.dcrf")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def _have(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install_shims():
    """Idempotent.  Real packages win when they are importable."""
    if not _have("timm"):
        timm = _mod("timm")
        data, models = _mod("timm.data"), _mod("timm.models")
        helpers, layers, registry = _mod("timm.models.helpers"), _mod("timm.models.layers"), _mod("timm.models.registry")
        timm.data, timm.models = data, models
        models.helpers, models.layers, models.registry = helpers, layers, registry
        data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
        data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
        helpers.load_pretrained = lambda *a, **k: None           # vit_* factories only (vit.py:1069-1211), unused

        class DropPath(nn.Module):                                # drop_path_rate is 0 everywhere (vit.py:151)
            def __init__(self, p=0.0):
                super().__init__()
                self.p = p

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
        layers.trunc_normal_ = lambda t, std=1.0, **k: nn.init.trunc_normal_(t, std=std, a=-2, b=2)
        models.resnet26d = models.resnet50d = None
        registry.register_model = lambda f: f
    if not _have("matplotlib"):
        mpl = _mod("matplotlib")
        plt = _mod("matplotlib.pyplot")
        mpl.pyplot = plt

        def get_cmap(name="jet"):                                 # utils/imutils.py:262 (tensorboard images only)
            def cmap(x):
                x = np.asarray(x, dtype=np.float32)
                return np.stack([x, 1.0 - np.abs(2.0 * x - 1.0), 1.0 - x, np.ones_like(x)], axis=-1)
            return cmap
        plt.get_cmap = get_cmap
    if not _have("tensorboardX"):
        tbx = _mod("tensorboardX")

        class SummaryWriter:                                      # train_final_voc.py:17,112: logging sink only
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None
        tbx.SummaryWriter = SummaryWriter
    if not _have("imageio"):
        from PIL import Image
        iio = _mod("imageio")
        v2 = _mod("imageio.v2")

        def imread(path):
            return np.asarray(Image.open(path))

        def imsave(path, arr):
            Image.fromarray(np.asarray(arr)).save(path)
        for m in (iio, v2):
            m.imread, m.imsave, m.imwrite = imread, imsave, imsave
        iio.v2 = v2
    if not _have("texttable"):
        tt = _mod("texttable")

        class Texttable:                                          # utils/pyutils.py:4,19-30: validation table
            def __init__(self):
                self.rows = []

            def header(self, h):
                self.rows.append(list(h))

            def add_row(self, r):
                self.rows.append(list(r))

            def draw(self):
                return "\n".join(" | ".join(f"{c:.3f}" if isinstance(c, float) else str(c) for c in r) for r in self.rows)
        tt.Texttable = Texttable
    try:
        # scikit-learn >= 1.6 returns a Python float from f1_score; the scripts call `.item()` on it
        # (train_final_voc.py:459-462 was written against scikit-learn 1.0.2, which returned numpy.float64)
        import sklearn.metrics as skm
        if not getattr(skm.f1_score, "_dupl_compat", False):
            _f1 = skm.f1_score

            def f1_score(*a, **k):
                return np.float64(_f1(*a, **k))
            f1_score._dupl_compat = True
            skm.f1_score = f1_score
    except ImportError:
        pass
    if not hasattr(np, "float"):
        np.float = float                                          # utils/optimizer.py:11-13 (CosWarmupAdamW, unused)


# ------------------------------------------------------------------------------------------------ offline "pretrained" weights
def deit_base_state_dict(seed=0):
    """A seeded stand-in for deit_base_patch16_224-b5f2ef4d.pth: same keys / shapes (timm DeiT-B/16: 152 tensors), values
    drawn like the reference's own init (vit.py:262-275) plus non-trivial biases / norms so that every path is exercised."""
    g = torch.Generator().manual_seed(1234 + seed)

    def tn(*s):
        return nn.init.trunc_normal_(torch.empty(*s), std=0.02, a=-2, b=2, generator=g)

    sd = {"cls_token": tn(1, 1, 768), "pos_embed": tn(1, 197, 768),
          "patch_embed.proj.weight": (torch.rand(768, 3, 16, 16, generator=g) - 0.5) * 2 / 768 ** 0.5,
          "patch_embed.proj.bias": (torch.rand(768, generator=g) - 0.5) * 2 / 768 ** 0.5}
    for i in range(12):
        b = f"blocks.{i}."
        for n, (o, k) in {"attn.qkv": (2304, 768), "attn.proj": (768, 768), "mlp.fc1": (3072, 768), "mlp.fc2": (768, 3072)}.items():
            sd[b + n + ".weight"] = tn(o, k)
            sd[b + n + ".bias"] = (torch.rand(o, generator=g) - 0.5) * 0.2 / k ** 0.5
        for n in ("norm1", "norm2"):
            sd[b + n + ".weight"] = 1 + 5 * tn(768)
            sd[b + n + ".bias"] = 5 * tn(768)
    sd["norm.weight"], sd["norm.bias"] = 1 + 5 * tn(768), 5 * tn(768)
    sd["head.weight"], sd["head.bias"] = tn(1000, 768), torch.zeros(1000)
    return sd


_orig_hub_load = None


def patch_hub_offline():
    """torch.hub.load_state_dict_from_url -> {"model": deit_base_state_dict()} for the DeiT URL when the checkpoint is not
    in the hub cache (no network in this image)."""
    global _orig_hub_load
    if _orig_hub_load is not None:
        return
    _orig_hub_load = torch.hub.load_state_dict_from_url

    def load(url, model_dir=None, *a, **k):
        fname = os.path.basename(url)
        if model_dir and os.path.exists(os.path.join(model_dir, fname)):
            return _orig_hub_load(url, model_dir, *a, **k)
        if "deit_base_patch16_224" in url:
            return {"model": deit_base_state_dict()}
        raise RuntimeError(f"offline: no cached checkpoint for {url}")
    torch.hub.load_state_dict_from_url = load


# ------------------------------------------------------------------------------------------------ importing the reference
class _Ns:
    pass


_cached = None


def _purge():
    saved = {}
    for k in list(sys.modules):
        if k.split(".")[0] in REF_PACKAGES:
            saved[k] = sys.modules.pop(k)
    return saved


def _bind_packages():
    """The reference's top-level directories have no __init__.py (namespace packages), and a REGULAR package of the same
    name elsewhere on sys.path would win (this image has HuggingFace `datasets` in site-packages): bind the names to the
    reference's directories explicitly."""
    for name in REF_PACKAGES:
        d = os.path.join(REF_ROOT, name)
        if os.path.isdir(d):
            pkg = types.ModuleType(name)
            pkg.__path__ = [d]
            pkg.__package__ = name
            sys.modules[name] = pkg


def load_reference():
    """-> namespace with the reference's own hot-path modules (model_dupl, PAR, losses, cam_helper, camutils, imutils,
    optimizer, train_helper is NOT imported: it pulls the dataset modules)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"no reference tree at {REF_ROOT}: run baseline/install_ref.sh in the authoring container")
    install_shims()
    saved = _purge()
    sys.path.insert(0, REF_ROOT)
    try:
        _bind_packages()
        ns = _Ns()
        ns.model_dupl = importlib.import_module("model.model_dupl")
        ns.PAR = importlib.import_module("model.PAR")
        ns.losses = importlib.import_module("model.losses")
        ns.cam_helper = importlib.import_module("utils.cam_helper")
        ns.camutils = importlib.import_module("utils.camutils")
        ns.imutils = importlib.import_module("utils.imutils")
        ns.optimizer = importlib.import_module("utils.optimizer")
        ns.evaluate = importlib.import_module("utils.evaluate")
    finally:
        sys.path.remove(REF_ROOT)
        _purge()
        sys.modules.update(saved)
    _cached = ns
    return ns


def enter_reference_tree(dropin=False):
    """For running the reference's scripts in this process: baseline/_ref first on sys.path (their `model`, `utils`,
    `datasets` packages become importable under those names) and cwd-independent.  dropin=True rebinds the six module
    names of DROPIN_MODULES to dupl_b200's implementations BEFORE the script imports them."""
    if not available():
        raise RuntimeError(f"no reference tree at {REF_ROOT}")
    install_shims()
    patch_hub_offline()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _purge()
    _bind_packages()
    if dropin:
        repo = os.path.dirname(HERE)
        if repo not in sys.path:
            sys.path.insert(1, repo)
        import dupl_b200.model.losses as d_losses
        import dupl_b200.model.model_dupl as d_model
        import dupl_b200.model.PAR as d_par
        import dupl_b200.utils.cam_helper as d_cam
        import dupl_b200.utils.camutils as d_camutils
        import dupl_b200.utils.dcrf as d_dcrf
        ref_model = importlib.import_module("model")       # the reference's packages (their __init__ import nothing heavy)
        ref_utils = importlib.import_module("utils")
        for pkg, name, mod in ((ref_model, "model_dupl", d_model), (ref_model, "PAR", d_par), (ref_model, "losses", d_losses),
                               (ref_utils, "cam_helper", d_cam), (ref_utils, "camutils", d_camutils), (ref_utils, "dcrf", d_dcrf)):
            sys.modules[f"{pkg.__name__}.{name}"] = mod
            setattr(pkg, name, mod)
