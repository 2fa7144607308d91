"""e2e of morsi_cuda_apply (pinned host buffers) for C2 under different chunk sizes: python scratch/e2e_sweep.py"""
import ctypes, os, sys, time
sys.path.insert(0, ".")
import imscript_b200 as M
from imscript_b200.binding import check
L = M.lib(); check(L.morsi_cuda_init(0))
w = h = 4096; planes = 3; n = w * h * planes
e = M.parse_element("disk7"); e_p = e.ctypes.data_as(M.binding._i32p)
hx, hy = M.binding._vp(), M.binding._vp()
check(L.morsi_cuda_host_alloc(ctypes.byref(hx), n * 4)); check(L.morsi_cuda_host_alloc(ctypes.byref(hy), n * 4))
d = M.DeviceBuffer(n * 4)
for p in range(planes): check(L.morsi_cuda_synth(d.ptr + p * w * h * 4, w, h, 0, p, 2, 0, None))
check(L.morsi_cuda_memcpy_d2h(hx, d.ptr, n * 4, None)); check(L.morsi_cuda_sync(None)); d.free()
ops = [M.OPS.index("opening"), M.OPS.index("closing")]
for mb in sys.argv[1:] or ["32"]:
    os.environ["MORSI_CUDA_CHUNK_MB"] = mb
    for o in ops: check(L.morsi_cuda_apply(o, e_p, hx, hy, w, h, planes))
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(5):
            for o in ops: check(L.morsi_cuda_apply(o, e_p, hx, hy, w, h, planes))
        best = min(best, (time.perf_counter() - t0) / 5)
    print(f"chunk {mb} MiB: {2 * n / best / 1e6:.0f} Mpixel/s ({best * 1e3:.2f} ms per step)", flush=True)
