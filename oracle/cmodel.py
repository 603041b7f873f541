"""ctypes face of oracle/omni_oracle.c -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference call site it restates; see omni_oracle.c for the arithmetic.
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        u8p, f32p, u16p, i32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint16),
                                 C.POINTER(C.c_int32))
        L.orc_resize_area_u8c3.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_resize_area_u8c3.restype = C.c_int
        L.orc_bgr2lab_u8.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_assign_f32.argtypes = [u8p, C.c_size_t, f32p, C.c_int, u8p]
        L.orc_assign_i16wrap.argtypes = [u8p, C.c_size_t, u8p, C.c_int, u8p]
        L.orc_onehot.argtypes = [u8p, C.c_size_t, u8p, C.c_int, u8p]
        L.orc_structuring_element.argtypes = [C.c_int, C.c_int, u8p]
        L.orc_morph_openclose.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.orc_gauss_weights.argtypes = [C.c_int, u16p]
        L.orc_gauss_weights.restype = C.c_int
        L.orc_gaussian_blur_u8.argtypes = [u8p, C.c_int, C.c_int, u8p, u16p, C.c_int]
        L.orc_canny_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_double, C.c_double, u8p, i32p, u8p]
        L.orc_layer_masks.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_edge_chain.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_double, C.c_double, u8p]
        L.orc_edge_chain.restype = C.c_int
        L.orc_thin_zhangsuen.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, i32p]
        L.orc_thin_zhangsuen.restype = C.c_int
        L.orc_swatch_masks.argtypes = [u8p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, u8p, i32p]
        L.orc_swatch_masks.restype = None
        _lib = L
    return _lib


def _p(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def resize_area(img, new_w, new_h):
    """01_resize.py:20 cv2.resize(img,(new_w,new_h),INTER_AREA) for HxWx3 u8, shrink only."""
    img = _u8(img)
    h, w, c = img.shape
    assert c == 3
    out = np.empty((new_h, new_w, 3), np.uint8)
    rc = lib().orc_resize_area_u8c3(_p(img), h, w, _p(out), new_h, new_w)
    if rc:
        raise ValueError("bad resize sizes")
    return out


def bgr2lab(img):
    """02_color_extract.py:35 cv2.cvtColor(img, COLOR_BGR2LAB), 8-bit."""
    img = _u8(img)
    out = np.empty_like(img)
    lib().orc_bgr2lab_u8(_p(img), img.size // 3, _p(out))
    return out


def assign_f32(px3, centers):
    """02_color_extract.py:53-55 nearest centre in float32 (first minimum)."""
    px3 = _u8(px3)
    ctr = np.ascontiguousarray(centers, dtype=np.float32)
    out = np.empty(px3.shape[:-1], np.uint8)
    lib().orc_assign_f32(_p(px3), px3.size // 3, _p(ctr, C.c_float), ctr.shape[0], _p(out))
    return out


def assign_i16wrap(img_rgb, palette_rgb):
    """process_colors.py:69-77 assign_labels incl. the int16 wrap of diff*diff."""
    img = _u8(img_rgb)
    pal = _u8(palette_rgb)
    out = np.empty(img.shape[:-1], np.uint8)
    lib().orc_assign_i16wrap(_p(img), img.size // 3, _p(pal), pal.shape[0], _p(out))
    return out


def structuring_element(shape, k):
    """cv2.getStructuringElement(shape,(k,k)); shape 0 = RECT, 2 = ELLIPSE."""
    se = np.empty((k, k), np.uint8)
    lib().orc_structuring_element(shape, k, _p(se))
    return se


def morph(img, se, op, iters):
    """cv2.morphologyEx(img, OPEN(op=0)|CLOSE(op=1), se, iterations=iters)."""
    out = _u8(img).copy()
    se = _u8(se)
    lib().orc_morph_openclose(_p(out), out.shape[0], out.shape[1], _p(se), se.shape[0], op, iters)
    return out


def gauss_weights(k):
    w = np.empty(k, np.uint16)
    if lib().orc_gauss_weights(k, _p(w, C.c_uint16)):
        raise ValueError(f"no fixed-point Gaussian table for k={k}")
    return w


def gaussian_blur(img, k):
    """03_edge_detect.py:33 cv2.GaussianBlur(img,(k,k),0) on u8."""
    img = _u8(img)
    out = np.empty_like(img)
    w = gauss_weights(k)
    lib().orc_gaussian_blur_u8(_p(img), img.shape[0], img.shape[1], _p(out), _p(w, C.c_uint16), k)
    return out


def canny(img, t1, t2, stages=False):
    """03_edge_detect.py:34 cv2.Canny(img, t1, t2) (aperture 3, L1)."""
    img = _u8(img)
    out = np.empty_like(img)
    if stages:
        mag = np.empty(img.shape, np.int32)
        nms = np.empty(img.shape, np.uint8)
        lib().orc_canny_u8(_p(img), img.shape[0], img.shape[1], t1, t2, _p(out), _p(mag, C.c_int32), _p(nms))
        return out, mag, nms
    lib().orc_canny_u8(_p(img), img.shape[0], img.shape[1], t1, t2, _p(out), None, None)
    return out


def layer_masks(labels, K, lut=None, open_iters=1, close_iters=1):
    """02_color_extract.py:146-154: K planes (lut[labels]==p)*255 -> RECT-3 open -> close."""
    labels = _u8(labels)
    h, w = labels.shape
    out = np.empty((K, h, w), np.uint8)
    lp = _p(_u8(lut)) if lut is not None else None
    lib().orc_layer_masks(_p(labels), h, w, lp, K, open_iters, close_iters, _p(out))
    return out


def edge_chain(mask, morph_k=3, open_iters=1, close_iters=1, ks=3, t1=50, t2=150):
    """03_edge_detect.py:23-34 for one layer mask."""
    mask = _u8(mask)
    out = np.empty_like(mask)
    rc = lib().orc_edge_chain(_p(mask), mask.shape[0], mask.shape[1], morph_k, open_iters, close_iters, ks,
                              float(t1), float(t2), _p(out))
    if rc:
        raise ValueError(f"edge_chain rc={rc}")
    return out


def thin_zhangsuen(img, max_iter=120, with_log=False):
    """04_find_contours.py:35-99 thinning_zhangsuen: u8 (>0 = foreground) -> skeleton {0,255}."""
    img = _u8(img)
    out = np.empty_like(img)
    removed = np.zeros(max(1, max_iter), np.int32)
    it = lib().orc_thin_zhangsuen(_p(img), img.shape[0], img.shape[1], _p(out), max_iter, _p(removed, C.c_int32))
    return (out, removed[:it].copy()) if with_log else out


def swatch_masks(img_bgr, colors, tol=30, with_choice=False):
    """02_color_extract.py:82-109 swatch branch: K masks (inRange of the better of RGB-reversed / as-is, RECT-3 open/close)."""
    img = _u8(img_bgr)
    col = np.ascontiguousarray(np.asarray(colors, np.int32).reshape(-1, 3))
    K = col.shape[0]
    out = np.empty((K,) + img.shape[:2], np.uint8)
    choice = np.zeros(K, np.int32)
    lib().orc_swatch_masks(_p(img), img.shape[0], img.shape[1], _p(col, C.c_int32), K, int(tol), _p(out), _p(choice, C.c_int32))
    return (out, choice) if with_choice else out
