"""CPU restatement (numpy) of the seven PIL operations of the reference's RandAugment (TEST INFRASTRUCTURE — never the
product).  utils/randomaug.py:161-262 (`augment_list`: AutoContrast, Equalize, Posterize, Color, Contrast, Brightness,
Sharpness) applied by utils/imutils.py:305-317 `augment_data_strong`.

The arithmetic lives in Pillow (third-party, C): ImageOps.autocontrast / equalize / posterize, ImageEnhance.{Color, Contrast,
Brightness, Sharpness} = Image.blend(degenerate, image, factor), ImageFilter.SMOOTH, Image.convert("L").  Pillow IS installed
in this image (12.2), so this restatement is pinned against Pillow itself, bit for bit, by tests/test_randaug_oracle.py; the
CUDA kernels (dupl_b200/csrc/augment.cu) are then compared with it and with Pillow on the GPU box.
All functions take and return uint8 arrays [H, W, 3].
"""
import numpy as np

OPS = ("AutoContrast", "Equalize", "Posterize", "Color", "Contrast", "Brightness", "Sharpness")
RANGES = ((0, 1), (0, 1), (0, 6), (0.1, 1.9), (0.1, 1.9), (0.1, 1.9), (0.1, 1.9))      # utils/randomaug.py:185-200


def magnitude(op_index, m):
    lo, hi = RANGES[op_index]
    return (float(m) / 30) * float(hi - lo) + lo                                      # utils/randomaug.py:262


def _lut_autocontrast(h):
    """ImageOps.autocontrast(cutoff=0) for one band's 256-bin histogram."""
    nz = np.nonzero(h)[0]
    lo, hi = int(nz[0]), int(nz[-1])
    lut = np.arange(256, dtype=np.int64)
    if hi <= lo:
        return lut.astype(np.uint8)
    scale = 255.0 / (hi - lo)
    offset = -lo * scale
    out = np.empty(256, np.int64)
    for ix in range(256):
        v = int(ix * scale + offset)
        out[ix] = 0 if v < 0 else (255 if v > 255 else v)
    return out.astype(np.uint8)


def _lut_equalize(h):
    """ImageOps.equalize for one band."""
    histo = [int(f) for f in h if f]
    if len(histo) <= 1:
        return np.arange(256, dtype=np.uint8)
    step = (sum(histo) - histo[-1]) // 255
    if not step:
        return np.arange(256, dtype=np.uint8)
    lut = np.empty(256, np.int64)
    n = step // 2
    for i in range(256):
        lut[i] = n // step
        n += int(h[i])
    return np.clip(lut, 0, 255).astype(np.uint8)                                        # Image.point() clips list entries (CLIP8)


def luma(img):
    """Image.convert("L") of an RGB image: (R*19595 + G*38470 + B*7471 + 0x8000) >> 16."""
    r, g, b = (img[..., c].astype(np.int64) for c in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(degenerate, img, alpha):
    """Image.blend(degenerate, img, alpha), 0 <= alpha <= 1: (UINT8)(d + alpha * (x - d)) in C float arithmetic (separately
    rounded multiply and add, truncation).  Outside [0, 1] Pillow clips first (same expression)."""
    a = np.float32(alpha)
    d = degenerate.astype(np.int32)
    x = img.astype(np.int32)
    t = (d.astype(np.float32) + (a * (x - d).astype(np.float32)).astype(np.float32)).astype(np.float32)
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def smooth(img):
    """image.filter(ImageFilter.SMOOTH): 3x3 kernel (1 1 1 / 1 5 1 / 1 1 1) / 13, border pixels copied; per pixel
    ss = 0.5 + (row y+1 triple) + (row y triple) + (row y-1 triple), each triple a*k0 + b*k1 + c*k2 in float, clip8 truncation."""
    k = (np.array([1, 1, 1, 1, 5, 1, 1, 1, 1], np.float64) / 13.0).astype(np.float32)
    H, W, _ = img.shape
    out = img.copy()
    f = img.astype(np.float32)

    def triple(row, k0, k1, k2):       # row: [h, W, 3] float32 rows aligned with the output rows
        return ((row[:, :-2] * k0).astype(np.float32) + (row[:, 1:-1] * k1).astype(np.float32)).astype(np.float32) + (row[:, 2:] * k2).astype(np.float32)

    ss = np.full((H - 2, W - 2, 3), 0.5, np.float32)
    ss = (ss + triple(f[2:], k[0], k[1], k[2]).astype(np.float32)).astype(np.float32)
    ss = (ss + triple(f[1:-1], k[3], k[4], k[5]).astype(np.float32)).astype(np.float32)
    ss = (ss + triple(f[:-2], k[6], k[7], k[8]).astype(np.float32)).astype(np.float32)
    out[1:-1, 1:-1] = np.where(ss <= 0, 0, np.where(ss >= 255, 255, ss.astype(np.int32))).astype(np.uint8)
    return out


def apply_op(img, op_index, val):
    name = OPS[op_index]
    if name in ("AutoContrast", "Equalize"):
        fn = _lut_autocontrast if name == "AutoContrast" else _lut_equalize
        out = np.empty_like(img)
        for c in range(3):
            out[..., c] = fn(np.bincount(img[..., c].ravel(), minlength=256))[img[..., c]]
        return out
    if name == "Posterize":
        bits = max(1, int(val))                                                         # utils/randomaug.py:107-110
        mask = ~(2 ** (8 - bits) - 1) & 0xFF
        return (img & np.uint8(mask)).astype(np.uint8)
    if name == "Color":
        deg = np.repeat(luma(img)[..., None], 3, axis=2)
    elif name == "Contrast":
        L = luma(img)
        mean = int(L.astype(np.int64).sum() / L.size + 0.5)                             # ImageStat.Stat(...).mean[0] + 0.5
        deg = np.full_like(img, mean)
    elif name == "Brightness":
        deg = np.zeros_like(img)
    else:
        deg = smooth(img)
    return blend(deg, img, val)


def augment_u8(img, op_indices, m):
    for k in op_indices:
        img = apply_op(img, k, magnitude(k, m))
    return img
