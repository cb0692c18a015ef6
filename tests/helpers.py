"""Shared test utilities: seeded synthetic inputs and deterministic model parameters."""
import numpy as np
import torch

MEAN = (123.675, 116.28, 103.53)
STD = (58.395, 57.12, 57.375)


def synth_images(b, h, w, seed=0):
    """Smoothed random uint8 images, ImageNet-normalised (SURVEY §8(d) synthetic inputs)."""
    g = torch.Generator().manual_seed(1000 + seed)
    u8 = torch.randint(0, 256, (b, 3, h, w), generator=g).float()
    u8 = torch.nn.functional.avg_pool2d(u8, 5, stride=1, padding=2, count_include_pad=False).round().clamp(0, 255)
    x = torch.empty_like(u8)
    for c in range(3):
        x[:, c] = (u8[:, c] - MEAN[c]) / STD[c]
    return x


def synth_cls_labels(b, K, seed=0):
    rng = np.random.RandomState(seed)
    out = np.zeros((b, K), np.float32)
    for i in range(b):
        n = rng.choice([1, 2, 3, 4], p=[0.6, 0.29, 0.09, 0.02])
        out[i, rng.choice(K, n, replace=False)] = 1
    return torch.from_numpy(out)


def synth_boxes(b, h, w, seed=0):
    rng = np.random.RandomState(seed)
    box = np.zeros((b, 4), np.int16)
    box[0] = [0, h, 0, w]
    for i in range(1, b):
        hh = rng.randint(h // 2, h + 1)
        ww = rng.randint(w // 2, w + 1)
        y0 = rng.randint(0, h - hh + 1)
        x0 = rng.randint(0, w - ww + 1)
        box[i] = [y0, y0 + hh, x0, x0 + ww]
    return torch.from_numpy(box)


def init_state_dict(num_classes=21, seed=0, students=(1, 2)):
    """Deterministic (CPU-generator) parameters under the reference's state-dict names, following the
    reference's initialisation scheme (vit.py:262-275: trunc-normal 0.02 linears, LayerNorm 1/0)."""
    g = torch.Generator().manual_seed(seed)
    K = num_classes - 1
    P = {}

    def tn(*s):
        return torch.nn.init.trunc_normal_(torch.empty(*s), std=0.02, a=-2, b=2, generator=g)

    def uni(*s, fan_in):
        return (torch.rand(*s, generator=g) - 0.5) * 2 / (fan_in ** 0.5)

    for br in students:
        e = f"branch{br}.encoder."
        P[e + "cls_token"] = tn(1, 1, 768)
        P[e + "pos_embed"] = tn(1, 197, 768)
        P[e + "patch_embed.proj.weight"] = uni(768, 3, 16, 16, fan_in=768)
        P[e + "patch_embed.proj.bias"] = uni(768, fan_in=768)
        for i in range(12):
            b = e + f"blocks.{i}."
            for n, (o, k) in {"attn.qkv": (2304, 768), "attn.proj": (768, 768), "mlp.fc1": (3072, 768),
                              "mlp.fc2": (768, 3072)}.items():
                P[b + n + ".weight"] = tn(o, k)
                P[b + n + ".bias"] = uni(o, fan_in=k) * 0.1
            for n in ("norm1", "norm2"):
                P[b + n + ".weight"] = 1 + 0.1 * tn(768) / 0.02
                P[b + n + ".bias"] = 0.1 * tn(768) / 0.02
        P[e + "norm.weight"] = 1 + 0.1 * tn(768) / 0.02
        P[e + "norm.bias"] = 0.1 * tn(768) / 0.02
        P[e + "head.weight"] = tn(1000, 768)
        P[e + "head.bias"] = torch.zeros(1000)
        d = f"branch{br}.decoder."
        P[d + "conv6.weight"] = uni(512, 768, 3, 3, fan_in=768 * 9)
        P[d + "conv7.weight"] = uni(512, 512, 3, 3, fan_in=512 * 9)
        P[d + "conv8.weight"] = uni(num_classes, 512, 1, 1, fan_in=512)
        for n in ("classifier", "aux_classifier"):
            P[f"branch{br}.{n}.weight"] = uni(K, 768, 1, 1, fan_in=768)
    return P


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def mscam_err(got, want, want_sum):
    """max over planes of |got - want| divided by the plane's conditioning 1/(range+1e-5), i.e. the error
    expressed in units of the UN-normalised CAM (see oracle.mscam_condition)."""
    from oracle import dupl_oracle as O
    e = (got.detach().cpu().float() - want).abs().amax((2, 3), keepdim=True) / O.mscam_condition(want_sum)
    return e.max().item()
