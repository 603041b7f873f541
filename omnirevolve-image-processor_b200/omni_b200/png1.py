"""1-bit greyscale PNG files for mask.png / edges.png.

The reference writes its {0,255} layers with cv2.imwrite as 8-bit PNGs (02_color_extract.py:156, 03_edge_detect.py:37) and every
consumer reads them back with cv2.imread(..., IMREAD_GRAYSCALE) (03:19, 04:215-216, ...).  A PNG of bit depth 1 decodes through
that call to the same uint8 array with values {0,255} (libpng scales 1-bit grey to 0 / 255), is 8 x less data to deflate, and
its scanline format -- leftmost pixel in the most significant bit -- is exactly OMNI_BITS_MSB_FIRST, so the packed planes that
come off the GPU are written without touching the pixels again.  OMNI_B200_PNG_BITS=8 in the environment keeps 8-bit files.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png1(bits: np.ndarray, w: int, level: int = 1) -> bytes:
    """bits: uint8 [H, ceil(w/8)], MSB-first rows (unused bits of the last byte 0).  Returns the PNG file contents."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    h, rb = bits.shape
    if rb != (w + 7) // 8:
        raise ValueError("row bytes do not match the width")
    raw = np.empty((h, rb + 1), np.uint8)            # filter type 0 in front of every scanline
    raw[:, 0] = 0
    raw[:, 1:] = bits
    ihdr = struct.pack(">IIBBBBB", w, h, 1, 0, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b"")


def write_png1(path: str, bits: np.ndarray, w: int) -> None:
    with open(path, "wb") as fh:
        fh.write(encode_png1(bits, w))


def use_1bit() -> bool:
    return os.environ.get("OMNI_B200_PNG_BITS", "1") != "8"


def unpack_rows(bits: np.ndarray, w: int) -> np.ndarray:
    """MSB-first packed rows -> uint8 {0,255} [H, w] (what cv2.imread of the 1-bit file returns)."""
    return np.unpackbits(np.ascontiguousarray(bits), axis=-1)[..., :w] * np.uint8(255)
