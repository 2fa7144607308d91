"""A/B of kernel variants (MORSI_CUDA_LIB=...): quick parity vs the oracle + device timings.
   python scratch/variant_check.py [tag]"""
import ctypes, os, sys, time
sys.path.insert(0, ".")
import numpy as np
import imscript_b200 as M
from imscript_b200.binding import check
from oracle import oracle
tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("MORSI_CUDA_LIB", "default")
L = M.lib(); o = oracle()
check(L.morsi_cuda_init(0))
def same(a, b):
    n = np.isnan(b)
    return np.array_equal(np.isnan(a), n) and np.array_equal(a.view(np.uint32)[~n], b.view(np.uint32)[~n])
bad = 0
for (h, w) in [(420, 600), (97, 132), (700, 1100)]:
    x = np.stack([M.synth_host(w, h, plane=p, seed=33, dist=2 if p == 1 else 0) for p in range(2)])
    x[x == 0] = 0.0
    for name in ("disk7", "disk15"):
        e = o.element(name)
        for op in ("erosion", "dilation", "opening", "closing", "tophat", "bothat", "gradient", "laplacian", "cblur", "oscillation", "enhance"):
            if not same(M.apply(op, e, x), o.apply(op, e, x)):
                bad += 1; print(tag, "MISMATCH", name, op, w, h)
def timeit(name, op, w, h, planes, steps=10):
    e = M.parse_element(name); e_p = e.ctypes.data_as(M.binding._i32p)
    n = w * h * planes
    dx, dy = M.DeviceBuffer(n * 4), M.DeviceBuffer(n * 4)
    for p in range(planes): check(L.morsi_cuda_synth(dx.ptr + p * w * h * 4, w, h, 0, p, 7, 0, None))
    opi = M.OPS.index(op)
    ev = [M.binding._vp() for _ in range(2)]
    for q in ev: check(L.morsi_cuda_event_create(ctypes.byref(q)))
    for _ in range(3): check(L.morsi_cuda_apply_device(opi, e_p, dx.ptr, dy.ptr, w, h, planes, None))
    check(L.morsi_cuda_sync(None))
    best = 1e9
    for rep in range(3):
        check(L.morsi_cuda_event_record(ev[0], None))
        for _ in range(steps): check(L.morsi_cuda_apply_device(opi, e_p, dx.ptr, dy.ptr, w, h, planes, None))
        check(L.morsi_cuda_event_record(ev[1], None)); check(L.morsi_cuda_sync(None))
        ms = ctypes.c_float(); check(L.morsi_cuda_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms)))
        best = min(best, ms.value / steps)
    dx.free(); dy.free()
    return best
res = []
for (name, op, w, h, pl, st) in [("disk7", "opening", 4096, 4096, 3, 20), ("disk7", "closing", 4096, 4096, 3, 20),
                                 ("disk7", "erosion", 4096, 4096, 3, 20), ("disk7", "tophat", 4096, 4096, 3, 20),
                                 ("disk7", "gradient", 4096, 4096, 3, 20), ("disk7", "cblur", 4096, 4096, 3, 20),
                                 ("disk7", "oscillation", 4096, 4096, 3, 20), ("disk15", "gradient", 40000, 10000, 1, 5),
                                 ("disk15", "tophat", 40000, 10000, 1, 5), ("disk15", "erosion", 40000, 10000, 1, 5)]:
    t = timeit(name, op, w, h, pl, st)
    res.append(f"{name} {op} {w}x{h}x{pl}: {t:.4f} ms ({8*w*h*pl/t/1e6:.0f} GB/s)")
print(f"[{tag}] parity {'OK' if not bad else 'BAD'} | " + " | ".join(res), flush=True)
