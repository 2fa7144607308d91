"""Parity of the CUDA path (through the C ABI) against the oracle and the
golden vectors recorded from the reference.  Bit-exact for everything; NaNs are
compared by class (payloads are platform noise)."""
import json
import os

import numpy as np
import pytest

import imscript_b200 as M
from oracle import OPS, oracle
from tests.golden.make_golden import adversarial_input, kat_input, sha

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    an, bn = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(an, bn) and \
        np.array_equal(a.view(np.uint32)[~an], b.view(np.uint32)[~bn])


def assert_same(got, want, what):
    if not same_bits(got, want):
        bad = np.argwhere(~((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))))
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} samples differ, first at {i}: got {got[i]!r} want {want[i]!r}")


def test_known_answers(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat_9_7.json")))
    gold_e = json.load(open(os.path.join(golden_dir, "elements.json")))
    x = kat_input()
    for key, (h, s, y00, ymid) in kat["cases"].items():
        name, op = key.split()
        y = M.apply(op, np.array(gold_e[name], dtype=np.int32), x)
        assert sha(y) == h, key


@pytest.mark.parametrize("path", [0, 1])
def test_adversarial_golden(golden_dir, path):
    z = np.load(os.path.join(golden_dir, "adversarial.npz"))
    x = z["x"]
    M.lib().morsi_cuda_set_path(path)
    try:
        for name in [k[2:] for k in z.files if k.startswith("e:")]:
            e = z["e:" + name]
            gold = z["y:" + name].view(np.float32)
            for k, op in enumerate(OPS):
                assert_same(M.apply(op, e, x), gold[k], f"{name} {op} path={path}")
    finally:
        M.lib().morsi_cuda_set_path(0)


SHAPES = [(1, 1), (1, 37), (41, 1), (2, 3), (7, 5), (33, 65), (64, 64), (97, 131), (130, 257)]
ELEMENTS = ["cross", "square", "disk2.5", "disk3", "disk4.2", "disk5", "disk7", "dysk4", "hrec6",
            "vrec3", "drec4", "Drec3", "hrec40", "vrec37"]


@pytest.mark.parametrize("dist", [0, 1, 2])
def test_vs_oracle_shapes_elements_ops(dist):
    """ragged / tiny / degenerate shapes x every element family x all 18 ops"""
    o = oracle()
    for si, (h, w) in enumerate(SHAPES):
        x = M.synth_host(w, h, seed=10 + si, dist=dist)
        for name in ELEMENTS:
            if h * w > 4000 and name in ("hrec40", "vrec37", "disk7") and dist == 1:
                continue
            e = o.element(name)
            for op in OPS:
                if op in ("median", "rank") and h * w > 9000 and e[0] > 60:
                    continue
                assert_same(M.apply(op, e, x), o.apply(op, e, x), f"{name} {op} {w}x{h} dist={dist}")


DISKS = ["disk2.5", "disk3", "disk3.5", "disk4", "disk4.2", "disk5", "disk5.1", "disk6", "disk7", "disk8",
         "disk9", "disk10", "disk11", "disk12", "disk13", "disk14", "disk15"]


@pytest.mark.parametrize("warps", ["2", "4"])
def test_disk_kernels_multi_strip_multi_band(warps, monkeypatch):
    """every compiled disk shape of k_disk.cu on an image wide and tall enough for
    several column strips and row bands per plane; both CTA sizes; the min/max
    operations incl. the fused two-stage ones; NaN / Inf / +-0 sprinkled in plane 1"""
    monkeypatch.setenv("MORSI_DISK_W", warps)
    o = oracle()
    h, w = 420, 1100 if warps == "4" else 600
    x = np.stack([M.synth_host(w, h, plane=p, seed=33, dist=2 if p == 1 else 0) for p in range(2)])
    x[0, 100:180, 200:330] = np.nan          # all-NaN windows for the small disks
    x[x == 0] = 0.0                          # no -0.0: the result must come from the fast kernels
    assert not np.any(x.view(np.uint32) == 0x80000000)
    xz = M.synth_host(w, 150, seed=34, dist=2)   # with -0.0: flag raised, order-preserving re-run
    for name in ("disk5", "disk7"):
        for op in ("opening", "dilation"):
            assert_same(M.apply(op, o.element(name), xz), o.apply(op, o.element(name), xz), f"{name} {op} -0")
    for name in DISKS:
        e = o.element(name)
        ops = ["erosion", "dilation", "opening", "closing", "tophat", "bothat"]
        if name in ("disk4.2", "disk7", "disk15"):
            ops += ["gradient", "oscillation", "cblur", "igradient"]
        for op in ops:
            assert_same(M.apply(op, e, x), o.apply(op, e, x), f"{name} {op} {w}x{h} W={warps}")


def test_disk_kernels_unaligned_width():
    """widths that are not a multiple of 4 (a TMA tensor map cannot address them):
    k_disk runs on NaN-padded pitched copies; the pad columns must stay absent for
    the second pass of oscillation, and the result is bit-identical all the same"""
    o = oracle()
    for (h, w) in [(230, 1001), (97, 131), (64, 67), (300, 258)]:
        x = np.stack([M.synth_host(w, h, plane=p, seed=77, dist=2 if p == 1 else 0) for p in range(2)])
        x[x == 0] = 0.0                      # no -0.0: the result must come from the fast kernels
        for name in ("disk4.2", "disk7", "disk15"):
            e = o.element(name)
            for op in ("erosion", "dilation", "opening", "closing", "tophat", "bothat", "gradient",
                       "oscillation", "laplacian", "cblur", "eblur"):
                assert_same(M.apply(op, e, x), o.apply(op, e, x), f"{name} {op} {w}x{h} unaligned")


@pytest.mark.parametrize("quad", ["1", "0"])
def test_median_fast_kernels(quad, monkeypatch):
    """median by every disk the shared-window kernel (k_median_quad) and the
    per-pixel kernel (k_median_fast) are compiled for: odd and even image sizes
    (partial 2x2 blocks at the right / bottom edge), heavy ties, a NaN/Inf
    plane (per-pixel general path), and the small 3x3 elements"""
    monkeypatch.setenv("MORSI_MEDIAN_QUAD", quad)
    o = oracle()
    for (h, w) in [(151, 203), (64, 130), (37, 66), (75, 260)]:     # 260: the 16-byte aligned 3x3 median kernel
        x = np.stack([M.synth_host(w, h, plane=p, seed=51, dist=p) for p in range(3)])
        x[x == 0] = 0.0
        for name in ["disk2.5", "disk3", "disk3.5", "disk4", "disk4.2", "disk5", "disk5.1", "disk6", "disk7",
                     "cross", "square"]:
            if quad == "0" and name in ("disk5.1", "disk6", "disk7") and h > 64:
                continue                     # without the quad kernel these take the (slow) exact path
            e = o.element(name)
            assert_same(M.apply("median", e, x), o.apply("median", e, x), f"{name} median {w}x{h} quad={quad}")


def test_runtime_rowrun_shapes():
    """k_runs: row-run elements whose shape is only known at run time -- disks outside the compiled table
    (disk6.5, disk16 ... disk32), rectangles given as user lists, medium hrec -- on aligned and odd widths,
    several strips / bands, NaN / Inf / +-0 in plane 1, and as row bands"""
    o = oracle()
    rect = np.array([35, 0, 0, 0] + [v for dx in range(-3, 4) for dy in range(-2, 3) for v in (dx, dy)], dtype=np.int32)   # 7x5
    cases = [("disk6.5", 203, 150), ("disk16", 260, 140), ("disk20", 300, 101), ("disk25.3", 131, 120),
             ("disk32", 120, 90), ("hrec13", 333, 40), ("hrec31", 512, 33), (rect, 97, 131), ("disk16", 1100, 300)]
    for name, w, h in cases:
        e = o.element(name) if isinstance(name, str) else name
        assert M.describe_element(e).startswith("rowrun"), name
        x = np.stack([M.synth_host(w, h, plane=p, seed=91, dist=2 if p == 1 else 0) for p in range(2)])
        x[0][x[0] == 0] = 0.0
        big = e[0] > 1500
        ops = ["erosion", "dilation", "opening", "tophat"] if big else \
            ["erosion", "dilation", "opening", "closing", "tophat", "bothat", "gradient", "igradient", "laplacian",
             "oscillation", "cblur", "eblur"]
        if w * h > 100000:
            ops = ["closing", "gradient"]
        for op in ops:
            assert_same(M.apply(op, e, x), o.apply(op, e, x), f"runs {name if isinstance(name, str) else 'rect7x5'} {op} {w}x{h}")
    # row bands through the band entry point
    e = o.element("disk16")
    x = M.synth_host(260, 200, seed=92, dist=2)
    want = o.apply("tophat", e, x)
    up, down = M.halo_rows("tophat", e)
    for b0, b1 in [(0, 70), (70, 71), (71, 200)]:
        i0, i1 = max(0, b0 - up), min(200, b1 + down)
        bx = M.DeviceBuffer.from_host(np.ascontiguousarray(x[i0:i1]))
        by = M.DeviceBuffer((b1 - b0) * 260 * 4)
        M.apply_band_device("tophat", e, bx, i0, i1 - i0, by, b0, b1 - b0, 260, 200)
        assert_same(by.to_host((b1 - b0, 260)), want[b0:b1], f"runs band [{b0},{b1})")


def test_long_lines_van_herk():
    """hrec / vrec long enough for the van Herk / Gil-Werman kernels (k_line.cuh), also as a
    user list with an off-centre line, on ragged sizes with NaN / Inf / +-0 sprinkled in"""
    o = oracle()
    for (h, w) in [(120, 333), (401, 96), (64, 700)]:
        x = np.stack([M.synth_host(w, h, plane=p, seed=81, dist=2 if p == 1 else 0) for p in range(2)])
        x[0][x[0] == 0] = 0.0
        shifted = np.array([100, 0, 0, 0] + [v for k in range(100) for v in (k - 30, 2)], dtype=np.int32)   # a row line at dy = 2
        for e in (o.element("hrec50"), o.element("vrec60"), o.element("hrec150"), shifted):
            for op in ("erosion", "dilation", "opening", "tophat", "gradient", "oscillation", "cblur"):
                assert_same(M.apply(op, e, x), o.apply(op, e, x), f"line n={e[0]} {op} {w}x{h}")


def test_all_nan_windows_and_constant_images():
    """windows with no usable neighbour give +-INF (src/morsi.c:63,77), also on the fast paths"""
    o = oracle()
    x = M.synth_host(64, 64, seed=21, dist=0)
    x[10:50, 8:56] = np.nan
    x[55:, :] = np.inf
    x[:, 60:] = -np.inf
    for name in ["cross", "square", "disk5", "disk7", "dysk4"]:
        e = o.element(name)
        for op in OPS:
            assert_same(M.apply(op, e, x), o.apply(op, e, x), f"nan block {name} {op}")
    z = np.zeros((40, 64), np.float32)
    for name in ["square", "disk7"]:
        for op in OPS:
            assert_same(M.apply(op, o.element(name), z), o.apply(op, o.element(name), z), f"zeros {name} {op}")


def test_multi_plane_and_user_lists():
    o = oracle()
    x = np.stack([M.synth_host(50, 40, plane=p, seed=4, dist=2 if p == 1 else 0) for p in range(3)])
    for e in ([4, 0, 1, -1, 0, 0, 2, 1, -1, 0, 3, -2], [6, 0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 2, 0, 2, -2, -1],
              [1, 0, 0, 0, 2, 1], [0, 0, 0, 0], o.element("disk3")):
        e = np.array(e, dtype=np.int32)
        for op in OPS:
            assert_same(M.apply(op, e, x), o.apply(op, e, x), f"user {list(e[:6])} {op}")


def test_device_and_band_entry_points():
    """1-device result == the same image processed as row bands"""
    o = oracle()
    h, w = 150, 70
    x = M.synth_host(w, h, seed=9, dist=2)
    dx = M.DeviceBuffer.from_host(x)
    dy = M.DeviceBuffer(x.nbytes)
    for name, op in [("disk5", "tophat"), ("cross", "gradient"), ("disk3", "median"), ("dysk4", "oscillation"),
                     ("disk7", "closing"), ("square", "rank"), ("vrec9", "opening")]:
        e = o.element(name)
        want = o.apply(op, e, x)
        M.apply_device(op, e, dx, dy, w, h)
        assert_same(dy.to_host((h, w)), want, f"device {name} {op}")
        up, down = M.halo_rows(op, e)
        got = np.empty_like(x)
        for b0, b1 in [(0, 40), (40, 41), (41, 110), (110, 150)]:
            i0, i1 = max(0, b0 - up), min(h, b1 + down)
            bx = M.DeviceBuffer.from_host(x[i0:i1])
            by = M.DeviceBuffer((b1 - b0) * w * 4)
            M.apply_band_device(op, e, bx, i0, i1 - i0, by, b0, b1 - b0, w, h)
            got[b0:b1] = by.to_host((b1 - b0, w))
        assert_same(got, want, f"bands {name} {op}")
    with pytest.raises(M.MorsiError):   # halo rows missing
        M.apply_band_device("tophat", "disk5", dx, 10, 20, dy, 10, 20, w, h)


def test_reference_signature_mirror():
    o = oracle()
    w, h = 31, 17
    x = M.synth_host(w, h, seed=2)
    e = M.morsi.build_disk(5.1)                       # corrview.c:37
    y = np.empty(w * h, np.float32)
    M.morsi.morsi_bothat(y, x.reshape(-1), w, h, e)   # corrview.c:38
    assert_same(y.reshape(h, w), o.apply("bothat", e, x), "morsi_bothat")
    outs = [np.empty(w * h, np.float32) if k % 2 == 0 else None for k in range(12)]
    M.morsi.morsi_all(*outs, x.reshape(-1), w, h, e)
    for out, op in zip(outs, ["erosion", "dilation", "opening", "closing", "gradient", "igradient",
                              "egradient", "laplacian", "enhance", "oscillation", "tophat", "bothat"]):
        if out is not None:
            assert_same(out.reshape(h, w), o.apply(op, e, x), "morsi_all " + op)


def test_apply_all_outputs():
    """morsi_cuda_apply_all (src/morsi.c:278-310): 12 outputs of one upload, several planes, some skipped"""
    import ctypes
    o = oracle()
    names = ["erosion", "dilation", "opening", "closing", "gradient", "igradient",
             "egradient", "laplacian", "enhance", "oscillation", "tophat", "bothat"]
    h, w = 150, 203
    x = np.stack([M.synth_host(w, h, plane=p, seed=61, dist=2 if p == 1 else 0) for p in range(2)])
    for ename, skip in [("disk7", ()), ("cross", (1, 6)), ("dysk4", (0, 2, 4, 8, 10))]:
        e = o.element(ename)
        outs = [None if k in skip else np.empty_like(x) for k in range(12)]
        ptrs = (ctypes.c_void_p * 12)()
        for k, out in enumerate(outs):
            if out is not None:
                ptrs[k] = out.ctypes.data
        ee = np.ascontiguousarray(e, dtype=np.int32)
        M.binding.check(M.lib().morsi_cuda_apply_all(ee.ctypes.data_as(M.binding._i32p), x.ctypes.data, ptrs, w, h, 2))
        for k, out in enumerate(outs):
            if out is not None:
                assert_same(out, o.apply(names[k], e, x), f"apply_all {ename} {names[k]}")


def test_full_size_properties():
    """BASELINE config sizes through size-independent properties (the oracle
    would need hours): duality, extensivity, idempotence, linearity under
    monotone maps, crop-vs-oracle spot checks."""
    o = oracle()
    w = h = 4096
    x = M.synth_host(w, h, seed=2, dist=0)
    e = o.element("disk7")
    ope = M.apply("opening", e, x)
    clo = M.apply("closing", e, x)
    assert np.all(ope <= x) and np.all(clo >= x)                       # anti-/extensive
    assert np.array_equal(M.apply("opening", e, ope), ope)             # idempotent
    assert np.array_equal(-M.apply("opening", e, -x), clo)             # duality (no zeros in -x? zeros are rare)
    assert np.array_equal(M.apply("tophat", e, x), x - ope)
    assert np.array_equal(M.apply("bothat", e, x), clo - x)
    for (r0, c0) in [(0, 0), (2000, 1500), (h - 96, w - 96), (0, w - 96)]:
        # window = crop grown by 2*reach where the image allows: artificial window
        # edges are then >= 2*reach away from the crop, real image edges coincide
        y0, x0 = max(0, r0 - 12), max(0, c0 - 12)
        win = x[y0:min(h, r0 + 96 + 12), x0:min(w, c0 + 96 + 12)]
        want = o.apply("opening", e, win)[r0 - y0:r0 - y0 + 96, c0 - x0:c0 - x0 + 96]
        assert_same(ope[r0:r0 + 96, c0:c0 + 96], want, f"opening crop at {(r0, c0)}")
