"""NumPy restatement of the Camera ctor's pyramid (TMVS/mvs/camera.cpp:63-92): cv::resize(INTER_AREA) of level 0 and the
Sobel(ksize=1) edge image. TEST INFRASTRUCTURE ONLY — the comparator of the device pyramid kernels (csrc/pmvs_pyramid.cuh);
itself pinned bit for bit against OpenCV (cv2) in tests/test_oracle_cpu.py. OpenCV is not under the reference tree; the
algorithm restated is imgproc's resize.cpp: computeResizeAreaTab + ResizeArea_Invoker for fractional scales,
ResizeAreaFast for integer scales (2x2: (sum + 2) >> 2; n x n: saturate_cast<uchar>(sum * (1.f / area)))."""
import math

import numpy as np


def _area_tab(ssize, dsize, scale):
    """OpenCV's INTER_AREA table for one axis and a non-integer scale (imgproc computeResizeAreaTab): per destination
    index the first source index, the tap count and the float32 weights."""
    max_taps = int(math.ceil(scale)) + 2
    start = np.zeros(dsize, dtype=np.int64)
    count = np.zeros(dsize, dtype=np.int64)
    W = np.zeros((dsize, max_taps), dtype=np.float32)
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = int(math.ceil(fsx1))
        sx2 = int(math.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        n, first = 0, sx1
        if sx1 - fsx1 > 1e-3:
            first = sx1 - 1
            W[dx, n] = (sx1 - fsx1) / cell
            n += 1
        for sx in range(sx1, sx2):
            W[dx, n] = 1.0 / cell
            n += 1
        if fsx2 - sx2 > 1e-3:
            W[dx, n] = min(min(fsx2 - sx2, 1.0), cell) / cell
            n += 1
        start[dx], count[dx] = first, n
    return start, count, W


def resize_area_fast(img, dcols, drows, iscale):
    """ResizeAreaFast (integer scale): block sums in int; 2x2 blocks -> (sum + 2) >> 2, else sum * (1.f / area) rounded half to
    even; blocks cut by the right / bottom border -> (float)sum / count."""
    rows, cols = img.shape
    out = np.zeros((drows, dcols), dtype=np.uint8)
    src = img.astype(np.int64)
    scale = np.float32(1.0) / np.float32(iscale * iscale)
    for dy in range(drows):
        y0, y1 = dy * iscale, min(dy * iscale + iscale, rows)
        for dx in range(dcols):
            x0, x1 = dx * iscale, min(dx * iscale + iscale, cols)
            blk = src[y0:y1, x0:x1]
            sm, cnt = int(blk.sum()), blk.size
            if cnt == iscale * iscale:
                v = (sm + 2) >> 2 if iscale == 2 else int(np.rint(np.float32(sm) * scale))
            else:
                v = int(np.rint(np.float32(sm) / np.float32(cnt))) if cnt else 0
            out[dy, dx] = min(max(v, 0), 255)
    return out


def resize_area(img, f):
    """cv::resize(img, Size(), f, f, INTER_AREA) for u8 single-channel, f < 1 (camera.cpp:85). Fractional scale: float32
    weighted horizontal sums per source row, then float32 weighted vertical sums, each accumulated in source order (the
    order csrc/pmvs_pyramid.cuh uses), round half to even, saturate. Integer scale: resize_area_fast."""
    rows, cols = img.shape
    dcols, drows = int(np.rint(cols * f)), int(np.rint(rows * f))
    scale = 1.0 / f
    iscale = int(np.rint(scale))
    if abs(scale - iscale) < np.finfo(np.float64).eps:
        return resize_area_fast(img, dcols, drows, iscale)
    xs, xn, Wx = _area_tab(cols, dcols, scale)
    ys, yn, Wy = _area_tab(rows, drows, scale)
    src = img.astype(np.float32)
    tmp = np.zeros((rows, dcols), dtype=np.float32)
    for k in range(Wx.shape[1]):
        m = k < xn
        if not m.any():
            break
        tmp[:, m] = tmp[:, m] + Wx[m, k][None, :] * src[:, xs[m] + k]
    out = np.zeros((drows, dcols), dtype=np.float32)
    for k in range(Wy.shape[1]):
        m = k < yn
        if not m.any():
            break
        out[m, :] = out[m, :] + Wy[m, k][:, None] * tmp[ys[m] + k, :]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def edge_image(grey):
    """Sobel(ksize=1) gradient magnitude, min-max normalised (camera.cpp:71-78, 87-91). ksize=1 is the
    [-1,0,1] central difference; the border is BORDER_REFLECT_101."""
    g = np.pad(grey.astype(np.float64), 1, mode="reflect")
    gx = g[1:-1, 2:] - g[1:-1, :-2]
    gy = g[2:, 1:-1] - g[:-2, 1:-1]
    e = np.sqrt(gx * gx + gy * gy)
    mn, mx = e.min(), e.max()
    return np.ascontiguousarray((e - mn) / (mx - mn))




def build_pyramid(grey0, lod_ratio, max_lod, with_edge):
    """Level list [(grey u8, edge f64|None)], levels 0..max_lod."""
    levels = []
    for i in range(max_lod + 1):
        g = np.ascontiguousarray(grey0) if i == 0 else resize_area(grey0, lod_ratio ** i)
        levels.append((g, edge_image(g) if with_edge else None))
    return levels
