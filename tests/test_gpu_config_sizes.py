"""Parity at the size of every BASELINE.json config (C1..C5), through the C ABI on
device-resident data -- the entry points bench.py times.  The oracle cannot run a
40000^2 image, so the large configs are compared on CROPS: the synthetic image is a
function of (seed, plane, row, column), so any window of it can be generated on the
host, pushed through the oracle with a margin of stages x reach (real image edges
coincide, artificial window edges stay outside the compared crop), and compared bit
for bit with the same crop of the device result.  Crops sit at the image corners,
the last rows / columns, interior positions, band seams and the last plane.

Reference semantics: src/morsi.c:30-35 (border rule), :65 (reach per stage),
:229-236 (tophat), :103-120 (median)."""
import ctypes

import numpy as np
import pytest

import imscript_b200 as M
from imscript_b200.binding import check
from oracle import oracle
from tests.test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu

CROP = 96


def window(w, h, r0, c0, margin, plane, seed, dist=0, ch=CROP, cw=CROP):
    """(rows y0:y1, cols x0:x1) of the synthetic image around the crop [r0,r0+ch) x [c0,c0+cw)"""
    y0, y1 = max(0, r0 - margin), min(h, r0 + ch + margin)
    x0, x1 = max(0, c0 - margin), min(w, c0 + cw + margin)
    rows = M.synth_host(w, y1 - y0, row0=y0, plane=plane, seed=seed, dist=dist)
    return np.ascontiguousarray(rows[:, x0:x1]), y0, x0


def read_rows(d_ptr, w, r_first, nrows):
    out = np.empty((nrows, w), np.float32)
    check(M.lib().morsi_cuda_memcpy_d2h(out.ctypes.data, d_ptr + r_first * w * 4, out.nbytes, None))
    check(M.lib().morsi_cuda_sync(None))
    return out


def check_crops(name, op, w, h, plane_index, seed, d_y_plane_ptr, y_row0, positions, what, dist=0):
    """d_y_plane_ptr: device pointer of output row y_row0 of the plane"""
    o = oracle()
    e = o.element(name)
    up, down = M.halo_rows(op, e)
    margin = max(up, down)
    for (r0, c0) in positions:
        win, y0, x0 = window(w, h, r0, c0, margin, plane_index, seed, dist)
        ch, cw = min(CROP, h - r0), min(CROP, w - c0)
        want = o.apply(op, e, win)[r0 - y0:r0 - y0 + ch, c0 - x0:c0 - x0 + cw]
        got = read_rows(d_y_plane_ptr, w, r0 - y_row0, ch)[:, c0:c0 + cw]
        assert_same(got, want, f"{what}: {name} {op} crop at row {r0}, column {c0}")


def run_device(name, op, w, h, planes, seed, plane0=0, dist=0):
    L = M.lib()
    n = w * h * planes
    d_x, d_y = M.DeviceBuffer(n * 4), M.DeviceBuffer(n * 4)
    for p in range(planes):
        check(L.morsi_cuda_synth(d_x.ptr + p * w * h * 4, w, h, 0, plane0 + p, seed, dist, None))
    M.apply_device(op, M.parse_element(name), d_x, d_y, w, h, planes)
    return d_x, d_y


def test_c1_full_image():
    """configs[0]: square erosion 1024x1024, every sample against the oracle; and through the host entry point"""
    w = h = 1024
    o = oracle()
    e = o.element("square")
    x = M.synth_host(w, h, seed=1)
    want = o.apply("erosion", e, x)
    d_x, d_y = run_device("square", "erosion", w, h, 1, seed=1)
    assert_same(d_y.to_host((h, w)), want, "C1 device")
    assert_same(M.apply("erosion", e, x), want, "C1 host")


def test_c2_crops_all_planes():
    """configs[1]: disk7 opening and closing 4096x4096x3 (device-resident, one launch over 3 planes)"""
    w = h = 4096
    pos = [(0, 0), (0, w - CROP), (h - CROP, 0), (h - CROP, w - CROP), (2011, 1501), (h - CROP, 2048 - 40)]
    for op in ("opening", "closing"):
        d_x, d_y = run_device("disk7", op, w, h, 3, seed=2)
        for p in range(3):
            check_crops("disk7", op, w, h, p, 2, d_y.ptr + p * w * h * 4, 0, pos[p::3] + pos[:1], f"C2 plane {p}")
        d_x.free(); d_y.free()


def test_c3_median_crops():
    """configs[2]: disk5 median 8192x8192 (k_median_quad tiles + the border kernel)"""
    w = h = 8192
    d_x, d_y = run_device("disk5", "median", w, h, 1, seed=3)
    pos = [(0, 0), (0, w - CROP), (h - CROP, 0), (h - CROP, w - CROP), (4096 - 48, 4096 - 48), (5000, 8), (h - CROP, 3001), (17, w - CROP)]
    check_crops("disk5", "median", w, h, 0, 3, d_y.ptr, 0, pos, "C3")


def test_c5_frames_crops():
    """configs[4] (one step's chunk): cross gradient on 64 RGB 1920x1080 frames = 192 planes in one launch"""
    w, h, planes = 1920, 1080, 192
    d_x, d_y = run_device("cross", "gradient", w, h, planes, seed=5)
    pos = [(0, 0), (0, w - CROP), (h - CROP, 0), (h - CROP, w - CROP), (500, 900), (h - CROP, 1000)]
    for p in (0, 1, 95, 190, 191):
        check_crops("cross", "gradient", w, h, p, 5, d_y.ptr + p * w * h * 4, 0, pos, f"C5 plane {p}")
    # the frame seam: the last rows of plane 190 must not see the first rows of plane 191
    check_crops("cross", "gradient", w, h, 190, 5, d_y.ptr + 190 * w * h * 4, 0, [(h - CROP, 0)], "C5 seam")


C4_W = C4_H = 40000


def test_c4_whole_plane_crops():
    """configs[3] on one GPU: disk15 tophat of the whole 40000x40000 plane (6.4 GB in, 6.4 GB out,
    byte offsets beyond 2^32, a tensor map of 40000 rows), crops incl. the bottom-right corner"""
    w, h = C4_W, C4_H
    d_x, d_y = run_device("disk15", "tophat", w, h, 1, seed=4)
    pos = [(0, 0), (0, w - CROP), (h - CROP, 0), (h - CROP, w - CROP), (20000, 17000), (39990 - CROP, 31000),
           (26843, 21000),                      # byte offset 2^32 of the plane falls inside this crop's rows
           (h - CROP, 20000 - 48)]
    check_crops("disk15", "tophat", w, h, 0, 4, d_y.ptr, 0, pos, "C4 whole plane")


@pytest.mark.parametrize("band", [(0, 5000), (17500, 22500), (39000, 40000), (39936, 40000)])
def test_c4_bands_at_true_offsets(band):
    """configs[3] as the sharded ranks run it: morsi_cuda_apply_band_device on 40000-wide bands at their
    TRUE row offsets (first band, an interior band, the last rows), halo rows held by the caller"""
    w, h = C4_W, C4_H
    b0, b1 = band
    L = M.lib()
    e = M.parse_element("disk15")
    up, down = M.halo_rows("tophat", e)
    assert (up, down) == (28, 28)
    i0, i1 = max(0, b0 - up), min(h, b1 + down)
    d_x = M.DeviceBuffer((i1 - i0) * w * 4)
    d_y = M.DeviceBuffer((b1 - b0) * w * 4)
    check(L.morsi_cuda_synth(d_x.ptr, w, i1 - i0, i0, 0, 4, 0, None))
    M.apply_band_device("tophat", e, d_x, i0, i1 - i0, d_y, b0, b1 - b0, w, h)
    rows = b1 - b0
    pos = [(b0, 0), (b0, w - CROP), (b1 - min(CROP, rows), 0), (b1 - min(CROP, rows), w - CROP),
           (b0 + rows // 2 - min(CROP, rows) // 2, 20000 - 48)]
    check_crops("disk15", "tophat", w, h, 0, 4, d_y.ptr, b0, pos, f"C4 band [{b0},{b1})")


def test_c4_band_missing_halo_is_refused():
    """the band contract: a caller that does not hold the halo rows gets MORSI_ERR_INVALID, not garbage"""
    w, h = 4000, 40000
    d_x = M.DeviceBuffer(100 * w * 4)
    d_y = M.DeviceBuffer(100 * w * 4)
    with pytest.raises(M.MorsiError):
        M.apply_band_device("tophat", "disk15", d_x, 20000, 100, d_y, 20000, 100, w, h)
