"""ad-hoc device timing: python scratch/time_op.py element op w h planes dist [steps]"""
import sys, ctypes
sys.path.insert(0, ".")
import imscript_b200 as M
from imscript_b200.binding import check
L = M.lib()
name, op, w, h, planes, dist = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
steps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
check(L.morsi_cuda_init(0))
e = M.parse_element(name); e_p = e.ctypes.data_as(M.binding._i32p)
n = w * h * planes
dx, dy = M.DeviceBuffer(n * 4), M.DeviceBuffer(n * 4)
for p in range(planes):
    check(L.morsi_cuda_synth(dx.ptr + p * w * h * 4, w, h, 0, p, 7, dist, None))
opi = M.OPS.index(op)
ev = [M.binding._vp() for _ in range(2)]
for x in ev: check(L.morsi_cuda_event_create(ctypes.byref(x)))
for _ in range(3): check(L.morsi_cuda_apply_device(opi, e_p, dx.ptr, dy.ptr, w, h, planes, None))
check(L.morsi_cuda_sync(None))
check(L.morsi_cuda_event_record(ev[0], None))
for _ in range(steps): check(L.morsi_cuda_apply_device(opi, e_p, dx.ptr, dy.ptr, w, h, planes, None))
check(L.morsi_cuda_event_record(ev[1], None)); check(L.morsi_cuda_sync(None))
ms = ctypes.c_float(); check(L.morsi_cuda_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms)))
t = ms.value / steps
print(f"{name} {op} {w}x{h}x{planes} dist={dist}: {t:.4f} ms  {n/t/1e3:.0f} Msample/s  {8*n/t/1e6:.0f} GB/s")
