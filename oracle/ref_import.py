"""Import the UNMODIFIED reference (Wu0409/DuPL) on CPU for oracle pinning.

TEST INFRASTRUCTURE ONLY.  Used by tools/make_golden.py and by container-side
tests to check oracle/ against the reference itself.  /root/reference does not
exist on the GPU box, so nothing under `-m gpu`, smoke() or bench.py calls this.

The reference imports a few third-party modules that are absent from this image
(timm, matplotlib); they are only touched at import time,
so in-memory stand-ins are enough.
"""
import os
import sys
import types

import torch.nn as nn

REF_ROOT = os.environ.get("DUPL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


def install_shims():
    if "timm" not in sys.modules:
        timm = _mod("timm")
        data = _mod("timm.data")
        models = _mod("timm.models")
        helpers = _mod("timm.models.helpers")
        layers = _mod("timm.models.layers")
        registry = _mod("timm.models.registry")
        timm.data, timm.models = data, models
        models.helpers, models.layers, models.registry = helpers, layers, registry
        data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
        data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
        helpers.load_pretrained = lambda *a, **k: None

        class DropPath(nn.Module):
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
    if "matplotlib" not in sys.modules:
        mpl = _mod("matplotlib")
        mpl.pyplot = _mod("matplotlib.pyplot")


class _RefModules:
    pass


_cached = None


def load():
    """Returns a namespace with the reference's hot-path modules.

    The reference uses top-level package names `model` and `utils`; they are
    imported with REF_ROOT first on sys.path and then REMOVED from sys.modules
    again so they cannot shadow dupl_b200's own drop-in modules of the same name.
    """
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_shims()
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")
             or k == "utils" or k.startswith("utils.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        import importlib
        ns = _RefModules()
        ns.model_dupl = importlib.import_module("model.model_dupl")
        ns.PAR = importlib.import_module("model.PAR")
        ns.losses = importlib.import_module("model.losses")
        ns.cam_helper = importlib.import_module("utils.cam_helper")
        ns.camutils = importlib.import_module("utils.camutils")
        ns.imutils = importlib.import_module("utils.imutils")
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")
                  or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cached = ns
    return ns
