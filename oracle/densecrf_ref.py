"""ctypes wrapper of oracle/densecrf_ref.c (TEST INFRASTRUCTURE; parity unpinned — see the C file).
Mirrors utils/dcrf.py:42-69 (DenseCRF.__call__) on numpy arrays."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libdensecrf_ref.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE], check=True)
        lib = C.CDLL(path)
        fp, u8p, ip = C.POINTER(C.c_float), C.POINTER(C.c_ubyte), C.POINTER(C.c_int)
        lib.densecrf_ref_inference.argtypes = [u8p, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                               C.c_float, C.c_float, C.c_int, fp, ip]
        lib.densecrf_bruteforce_inference.argtypes = [u8p, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                      C.c_float, C.c_float, C.c_int, fp]
        _LIB = lib
    return _LIB


def unary_from_softmax(probs, clip=1e-5):
    """pydensecrf.utils.unary_from_softmax: -log(clip(p, clip, 1)) as float32 [C, H*W]."""
    c = probs.shape[0]
    return (-np.log(np.clip(probs, clip, 1.0))).reshape(c, -1).astype(np.float32)


def _run(fn, image, unary, H, W, pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std, iters, sizes=None):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    unary = np.ascontiguousarray(unary, dtype=np.float32)
    Cn = unary.shape[0]
    out = np.empty((Cn, H * W), np.float32)
    args = [image.ctypes.data_as(C.POINTER(C.c_ubyte)), unary.ctypes.data_as(C.POINTER(C.c_float)), W, H, Cn,
            pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std, iters, out.ctypes.data_as(C.POINTER(C.c_float))]
    if sizes is not None:
        args.append(sizes.ctypes.data_as(C.POINTER(C.c_int)))
    rc = fn(*args)
    assert rc == 0
    return out.reshape(Cn, H, W)


class DenseCRF:
    """Same constructor / call signature as utils/dcrf.py:42-69."""

    def __init__(self, iter_max, pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std):
        self.iter_max, self.pos_w, self.pos_xy_std = iter_max, pos_w, pos_xy_std
        self.bi_w, self.bi_xy_std, self.bi_rgb_std = bi_w, bi_xy_std, bi_rgb_std
        self.lattice_sizes = np.zeros(2, np.int32)

    def __call__(self, image, probmap):
        Cn, H, W = probmap.shape
        U = unary_from_softmax(probmap)
        return _run(_lib().densecrf_ref_inference, image, U, H, W, self.pos_w, self.pos_xy_std, self.bi_w, self.bi_xy_std,
                    self.bi_rgb_std, self.iter_max, self.lattice_sizes)

    def bruteforce(self, image, probmap):
        Cn, H, W = probmap.shape
        U = unary_from_softmax(probmap)
        return _run(_lib().densecrf_bruteforce_inference, image, U, H, W, self.pos_w, self.pos_xy_std, self.bi_w,
                    self.bi_xy_std, self.bi_rgb_std, self.iter_max)
