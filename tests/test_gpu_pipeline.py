"""The callers and data formats either side of the hot path and the
drop-in boundary exercised from C: morsi_all with shared passes (device and
host), pixel-interleaved images split on the device, the streaming entry point,
the `all` / streaming spellings of the CLI, a C program linked against
libmorsi_compat the way corrview.c would be, the multi-call (im.c) link, two
host threads on one device, and pageable (malloc'd) host buffers."""
import ctypes
import os
import subprocess
import threading

import numpy as np
import pytest

import imscript_b200 as M
from imscript_b200.binding import check
from oracle import OPS, oracle
from tests.test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "imscript_b200", "lib")
ALL_NAMES = ["erosion", "dilation", "opening", "closing", "gradient", "igradient",
             "egradient", "laplacian", "enhance", "oscillation", "tophat", "bothat"]


def ref_checker():
    """the compiled reference where it was shipped, else the restatement"""
    from oracle.oracle import Reference
    if os.path.exists(Reference.path):
        from oracle import reference
        return reference()
    return oracle()


def test_apply_all_device_shares_passes_and_matches():
    o = oracle()
    L = M.lib()
    h, w, planes = 140, 204, 2
    x = np.stack([M.synth_host(w, h, plane=p, seed=71, dist=2 if p == 1 else 0) for p in range(planes)])
    dx = M.DeviceBuffer.from_host(x)
    for ename, skip in [("disk7", ()), ("cross", (0, 1, 2, 3)), ("dysk4", (4, 5, 6, 7, 8)), ("disk5", (0, 2, 9, 10, 11))]:
        e = o.element(ename)
        bufs = [None if k in skip else M.DeviceBuffer(x.nbytes) for k in range(12)]
        ptrs = (ctypes.c_void_p * 12)()
        for k, b in enumerate(bufs):
            if b is not None:
                ptrs[k] = b.ptr
        ee = np.ascontiguousarray(e, dtype=np.int32)
        L.morsi_cuda_launch_count_reset()
        check(L.morsi_cuda_apply_all_device(ee.ctypes.data_as(M.binding._i32p), dx.ptr, ptrs, w, h, planes, None))
        check(L.morsi_cuda_sync(None))
        for k, b in enumerate(bufs):
            if b is not None:
                assert_same(b.to_host(x.shape), o.apply(ALL_NAMES[k], e, x), f"apply_all_device {ename} {ALL_NAMES[k]}")


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_interleaved_images_split_on_device(dtype, monkeypatch):
    """iio's vec layout (pixel-interleaved) in, vec floats out: = split, operate per plane, join"""
    o = oracle()
    rng = np.random.default_rng(5)
    for (h, w, pd, rows) in [(61, 83, 3, None), (130, 64, 4, "16"), (40, 37, 1, None), (75, 50, 2, "9")]:
        if rows:
            monkeypatch.setenv("MORSI_CUDA_CHUNK_ROWS", rows)
        else:
            monkeypatch.delenv("MORSI_CUDA_CHUNK_ROWS", raising=False)
        if dtype == np.float32:
            x = rng.random((h, w, pd), dtype=np.float32)
            x[3, 5, 0] = np.nan
        else:
            x = rng.integers(0, np.iinfo(dtype).max + 1, size=(h, w, pd)).astype(dtype)
        planes = np.ascontiguousarray(np.moveaxis(x.astype(np.float32), -1, 0))
        for name, op in [("disk5", "tophat"), ("cross", "gradient"), ("disk3", "median"), ("hrec4", "dilation")]:
            e = o.element(name)
            want = np.moveaxis(o.apply(op, e, planes), 0, -1)
            assert_same(M.apply_interleaved(op, e, x), want, f"interleaved {dtype.__name__} {name} {op} {w}x{h}x{pd}")


def test_stream_callbacks(monkeypatch):
    """row bands pulled and pushed through callbacks = the whole-image result; rows arrive in order"""
    o = oracle()
    h, w, planes = 333, 120, 2
    x = np.stack([M.synth_host(w, h, plane=p, seed=8, dist=2 if p == 1 else 0) for p in range(planes)])
    for rows in ("11", "64", None):
        if rows:
            monkeypatch.setenv("MORSI_CUDA_CHUNK_ROWS", rows)
        else:
            monkeypatch.delenv("MORSI_CUDA_CHUNK_ROWS", raising=False)
        for name, op in [("disk7", "closing"), ("square", "erosion"), ("disk4.2", "median"), ("vrec9", "oscillation")]:
            e = o.element(name)
            y = np.full_like(x, np.nan)
            order = []

            def rd(plane, row0, nrows):
                return x[plane, row0:row0 + nrows]

            def wr(plane, row0, rows_):
                order.append((plane, row0))
                y[plane, row0:row0 + rows_.shape[0]] = rows_
            M.apply_stream(op, e, w, h, planes, rd, wr)
            assert order == sorted(order)
            assert_same(y, o.apply(op, e, x), f"stream {name} {op} rows={rows}")
    with pytest.raises(M.MorsiError):
        M.apply_stream("erosion", "cross", w, h, 1, lambda *a: 1 / 0, lambda *a: None)


def test_cli_all_and_streaming(tmp_path):
    cli = os.path.join(LIBDIR, "morsi")
    x = M.synth_host(150, 90, seed=3)
    fin = str(tmp_path / "in.npy")
    np.save(fin, x)
    p = subprocess.run([cli, "disk4.2", "all", fin, str(tmp_path / "o_%s.npy")], capture_output=True)
    assert p.returncode == 0, p.stderr
    for name in ALL_NAMES:
        single = str(tmp_path / "single.npy")
        assert subprocess.run([cli, "disk4.2", name, fin, single]).returncode == 0
        assert open(single, "rb").read() == open(str(tmp_path / f"o_{name}.npy"), "rb").read(), name
    assert subprocess.run([cli, "disk4.2", "all", fin, str(tmp_path / "nopattern.npy")], capture_output=True).returncode == 1
    # streaming: .npy in, .npy out, the host never holds the image
    env = dict(os.environ, MORSI_CUDA_STREAM="1", MORSI_CUDA_CHUNK_ROWS="17")
    fs, fn = str(tmp_path / "stream.npy"), str(tmp_path / "normal.npy")
    assert subprocess.run([cli, "disk7", "tophat", fin, fs], env=env).returncode == 0
    assert subprocess.run([cli, "disk7", "tophat", fin, fn]).returncode == 0
    a, b = np.load(fs), np.load(fn)
    assert a.dtype == np.float32 and a.size == b.size
    assert_same(a.reshape(b.shape), b, "streamed CLI")
    assert subprocess.run([cli, "disk7", "tophat", fn, fs], env=env).returncode == 0     # iio's own (h, w, 1) header streams too
    assert np.load(fs).shape == b.shape
    # the multi-call form (src/im.c): `im morsi ...`
    im = os.path.join(LIBDIR, "im_like")
    fo = str(tmp_path / "im.npy")
    assert subprocess.run([im, "morsi", "disk7", "tophat", fin, fo]).returncode == 0
    assert open(fo, "rb").read() == open(fn, "rb").read()


def test_c_caller_linked_against_compat(tmp_path):
    """tests/c/compat_caller.c: build_disk(5.1), morsi_bothat, morsi_enhance, morsi_median, morsi_all called
    from C exactly like src/ftr/webcam/corrview.c does, linked -lmorsi_compat -lmorsi_cuda; results vs the reference"""
    r = ref_checker()
    o = oracle()
    w, h = 97, 61
    x = M.synth_host(w, h, seed=13, dist=2)
    fin = str(tmp_path / "x.raw")
    x.tofile(fin)
    prefix = str(tmp_path / "o_")
    p = subprocess.run([os.path.join(LIBDIR, "compat_caller"), str(w), str(h), fin, prefix], capture_output=True)
    assert p.returncode == 0, p.stderr
    e = o.element("disk5.1")
    for name in ["bothat", "enhance", "median"] + ["all_" + n for k, n in enumerate(ALL_NAMES) if k not in (1, 6)]:
        got = np.fromfile(prefix + name + ".raw", dtype=np.float32).reshape(h, w)
        assert_same(got, r.apply(name.replace("all_", ""), e, x), "C caller " + name)
    assert not os.path.exists(prefix + "all_dilation.raw")


def test_two_host_threads_one_device_and_pageable_buffers():
    """morsi_cuda_apply from two threads at once (ctypes drops the GIL) on malloc'd numpy buffers > 8 MiB:
    the calls are serialised per device and the buffers page-locked for the call"""
    o = oracle()
    w, h = 2048, 1100                       # 9 MB per plane: above the registration threshold
    xs = [M.synth_host(w, h, seed=21 + t) for t in range(2)]
    ops = [("disk7", "opening"), ("cross", "gradient")]
    out = [None, None]

    def work(t):
        for _ in range(3):
            out[t] = M.apply(ops[t][1], o.element(ops[t][0]), xs[t])
    th = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for t in range(2):
        e = o.element(ops[t][0])
        crop = o.apply(ops[t][1], e, xs[t][:80])
        assert_same(out[t][:60], crop[:60], f"thread {t}")
        assert_same(out[t], M.apply(ops[t][1], e, xs[t]), f"thread {t} repeat")


def qeasy_numpy(x, black, white, to_u8):
    """src/qeasy.c:57-69 restated: float arithmetic left to right, floor, saturate"""
    x = x.astype(np.float32)
    with np.errstate(all="ignore"):
        v = np.floor((np.float32(255) * (x - np.float32(black))) / (np.float32(white) - np.float32(black))).astype(np.float32)
        if not to_u8:
            return v
        g = np.where(np.isnan(v), 0, np.clip(v, 0, 255)).astype(np.uint8)
    return g


def test_pipe_chain_and_qeasy():
    """morsi | morsi | qeasy as one call: the operations back to back on the device, the quantiser last"""
    o = oracle()
    h, w = 120, 203
    x = np.stack([M.synth_host(w, h, plane=p, seed=17, dist=2 if p == 1 else 0) * 200 - 50 for p in range(2)]).astype(np.float32)
    steps = [("opening", o.element("disk3")), ("gradient", o.element("cross")), ("tophat", o.element("disk7"))]
    want = x
    for op, e in steps:
        want = o.apply(op, e, want)
    assert_same(M.apply_chain(steps, x), want, "chain of three operations")
    assert_same(M.apply_chain(steps[:1], x, quant=(40, -40, False)), qeasy_numpy(o.apply("opening", steps[0][1], x), 40, -40, False), "qeasy -f")
    for (b, wh) in [(0, 60), (40, -40), (-10.5, 3.25)]:
        lap = o.apply("laplacian", o.element("cross"), x)
        got = M.apply_chain([("laplacian", o.element("cross"))], x, quant=(b, wh, True))
        assert got.dtype == np.uint8 and np.array_equal(got, qeasy_numpy(lap, b, wh, True)), (b, wh)


def test_cli_qeasy_chain(tmp_path):
    """MORSI_CUDA_QEASY="0 60" morsi cross tophat in out.npy  ==  morsi cross tophat in | qeasy 0 60 - out"""
    cli = os.path.join(LIBDIR, "morsi")
    rng = np.random.default_rng(3)
    x = (rng.random((70, 90, 3), dtype=np.float32) * 255).astype(np.float32)
    fin, fout = str(tmp_path / "in.npy"), str(tmp_path / "out.npy")
    np.save(fin, x)
    env = dict(os.environ, MORSI_CUDA_QEASY="0 60")
    p = subprocess.run([cli, "cross", "tophat", fin, fout], env=env, capture_output=True)
    assert p.returncode == 0, p.stderr
    got = np.load(fout)
    o = oracle()
    planes = np.ascontiguousarray(np.moveaxis(x, -1, 0))
    want = np.moveaxis(qeasy_numpy(o.apply("tophat", o.element("cross"), planes), 0, 60, True), 0, -1)
    assert got.dtype == np.uint8 and np.array_equal(got.reshape(want.shape), want)
