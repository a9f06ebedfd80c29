"""TEST INFRASTRUCTURE (not shipped, not imported by deephumor_b200/): CPU restatement of the image resize that precedes
the hot path -- torchvision `transforms.Resize((224, 224))` applied to a PIL image (deephumor_demo.ipynb cell 11;
deephumor/data/datasets.py:48-53,94-98).

The arithmetic lives in a third-party, un-vendored dependency: torchvision (`F.resize` -> `PIL.Image.resize(size,
BILINEAR)`) -> Pillow's `ImagingResample` (src/libImaging/Resample.c; Pillow 12.2.0 and torchvision 0.26 in this image,
unpinned in the reference's requirements.txt).  This file restates Pillow's published algorithm for 8-bit RGB:
`precompute_coeffs` (triangle filter, support scaled by the reduction factor), `normalize_coeffs_8bpc` (22-bit fixed
point), `ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc` with a uint8 intermediate restricted to
the rows the vertical pass reads.  Pinned: tests/test_resize.py compares it with PIL.Image.resize / torchvision on
random and structured images (bit-exact), on this box and on the GPU box (both libraries ship in the image).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _tri(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size, out_size):
    """-> (bounds [out,2] int (first source index, tap count), taps [out,ksize] int32 fixed point)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    taps = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)              # int(): truncation toward zero, like the C cast
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_tri((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            taps[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, taps


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass_x(src, bounds, taps, out_w):
    """taps along axis 1 of src [R, W, 3] -> uint8 [R, out_w, 3]."""
    s64 = src.astype(np.int64)
    out = np.empty((src.shape[0], out_w, 3), dtype=np.uint8)
    for xx in range(out_w):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        out[:, xx, :] = _clip8((1 << (PRECISION_BITS - 1)) + np.tensordot(s64[:, x0:x0 + n, :], taps[xx, :n], axes=([1], [0])))
    return out


def _pass_y(src, bounds, taps, out_h, first=0):
    """taps along axis 0 of src [R, W, 3] (row 0 of src = source row `first`) -> uint8 [out_h, W, 3]."""
    s64 = src.astype(np.int64)
    out = np.empty((out_h, src.shape[1], 3), dtype=np.uint8)
    for yy in range(out_h):
        y0, n = int(bounds[yy, 0]) - first, int(bounds[yy, 1])
        out[yy] = _clip8((1 << (PRECISION_BITS - 1)) + np.tensordot(taps[yy, :n], s64[y0:y0 + n], axes=([0], [0])))
    return out


def resize_bilinear(img, out_h=224, out_w=224):
    """img uint8 [H,W,3] -> uint8 [out_h,out_w,3], equal to np.asarray(PIL.Image.fromarray(img).resize((out_w, out_h),
    BILINEAR)).  Pass order: horizontal then vertical (the horizontal pass restricted to the rows the vertical one reads),
    except for images taller than 100 x their width whose height is reduced (H > 100 W and H > out_h), where Pillow 12.2 is
    observed to run the vertical pass first (the uint8 intermediate makes the order visible in the last bit; pinned against
    PIL on both sides of the threshold by tests/test_resize.py)."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    H, W, _ = img.shape
    bh, kh = precompute_coeffs(W, out_w)
    bv, kv = precompute_coeffs(H, out_h)
    if H > 100 * W and H > out_h:
        return _pass_x(_pass_y(img, bv, kv, out_h), bh, kh, out_w)
    first = int(bv[0, 0])
    last = int(bv[-1, 0] + bv[-1, 1])
    return _pass_y(_pass_x(img[first:last], bh, kh, out_w), bv, kv, out_h, first)
