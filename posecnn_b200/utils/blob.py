"""Host-side halves of the input pre-processing.

The reference builds fp32 blobs on the CPU (`lib/fcn/test.py:37-110`: BGR - PIXEL_MEANS, depth
`clip(d / 2000, 0, 1) * 255` tiled to three channels minus PIXEL_MEANS, `lib/utils/blob.py:48-71`: pad to a multiple
of 16).  Here the colour path stays uint8 all the way to the GPU — the mean subtraction and the bf16 conversion happen
inside the conv1_1 loader (`k_conv1_tc`) — so the host only pads and stacks; the depth path (non-integer values) is
prepared as fp32.
"""
from __future__ import annotations

import numpy as np

PIXEL_MEANS = np.array([102.9801, 115.9465, 122.7717], dtype=np.float32)      # lib/fcn/config.py:242 (BGR)


def _padding(height: int, width: int, factor: int):
    return (-height) % factor, (-width) % factor


def pad_im(im: np.ndarray, factor: int, value=0) -> np.ndarray:
    """Pad bottom / right to the next multiple of `factor` (lib/utils/blob.py:48-58)."""
    ph, pw = _padding(im.shape[0], im.shape[1], factor)
    out = np.full((im.shape[0] + ph, im.shape[1] + pw) + im.shape[2:], value, dtype=im.dtype)
    out[:im.shape[0], :im.shape[1]] = im
    return out


def unpad_im(im: np.ndarray, factor: int) -> np.ndarray:
    """lib/utils/blob.py:61-71, literally: removes `ceil(h / factor) * factor - h` rows computed from the PADDED
    height -- zero for an image that pad_im produced (callers crop results with the original size instead)."""
    ph, pw = _padding(im.shape[0], im.shape[1], factor)
    return im[:im.shape[0] - ph, :im.shape[1] - pw]


def color_blob(ims, factor: int = 16) -> np.ndarray:
    """List of HxWx3 uint8 BGR images -> [N, Hp, Wp, 3] uint8, zero padded to the largest (padded) size
    (im_list_to_blob, blob.py:12-28).  NOTE: the reference pads AFTER the mean subtraction, i.e. with the value 0 of the
    mean-subtracted image; a uint8 blob cannot express that, so padded pixels carry the rounded PIXEL_MEANS, which the
    device-side subtraction maps to within 0.5 of zero.  640x480 inputs need no padding."""
    padded = [pad_im(np.ascontiguousarray(im, dtype=np.uint8), factor, value=0) for im in ims]
    hp, wp = max(p.shape[0] for p in padded), max(p.shape[1] for p in padded)
    fill = np.rint(PIXEL_MEANS).astype(np.uint8)
    blob = np.empty((len(ims), hp, wp, 3), np.uint8)
    blob[:] = fill
    for i, im in enumerate(ims):
        blob[i, :im.shape[0], :im.shape[1]] = im
    return blob


def depth_blob(depths, factor: int = 16) -> np.ndarray:
    """List of HxW depth images (raw sensor units) -> [N, Hp, Wp, 3] float32 = clip(d / 2000, 0, 1) * 255 tiled to three
    channels minus PIXEL_MEANS (test.py:72-76), zero padded (im_list_to_blob)."""
    out = []
    for d in depths:
        v = np.clip(np.asarray(d, dtype=np.float32) / np.float32(2000.0), 0, 1) * np.float32(255)
        out.append(pad_im(np.repeat(v[:, :, None], 3, axis=2) - PIXEL_MEANS, factor, value=0))
    hp, wp = max(p.shape[0] for p in out), max(p.shape[1] for p in out)
    blob = np.zeros((len(out), hp, wp, 3), np.float32)
    for i, p in enumerate(out):
        blob[i, :p.shape[0], :p.shape[1]] = p
    return blob
