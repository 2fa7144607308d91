"""Host-side mirror of the reference's operator interface (src/morsi.c:56-310):
one function per operation with the reference's names and argument order
``morsi_<op>(y, x, w, h, e)`` on caller-owned planar float32 buffers, plus
``build_disk``.  Each call goes through the C ABI (morsi_cuda_apply) to the
sm_100a kernels; errors raise MorsiError (the reference calls exit(-1))."""
import numpy as np

from . import binding as _b


def _run(op, y, x, w, h, e):
    x = np.asarray(x)
    y = np.asarray(y)
    if x.dtype != np.float32 or y.dtype != np.float32 or not y.flags.c_contiguous or not y.flags.writeable:
        raise _b.MorsiError(1, "x and y must be float32, y contiguous and writeable")
    if x.size != w * h or y.size != w * h:
        raise _b.MorsiError(1, "x and y must hold w*h samples")
    y.reshape(-1)[:] = _b.apply(op, e, x.reshape(h, w)).reshape(-1)


def _make(op):
    def f(y, x, w, h, e):
        _run(op, y, x, w, h, e)
    f.__name__ = "morsi_" + op
    f.__doc__ = f"GPU drop-in for morsi_{op}(float *y, float *x, int w, int h, int *e)."
    return f


for _op in _b.OPS:
    globals()["morsi_" + _op] = _make(_op)


def build_disk(radius):
    """src/morsi.c:313-330; None when radius <= 1."""
    return _b.build_element("disk", radius)


def morsi_all(o_ero, o_dil, o_ope, o_clo, o_grad, o_igrad, o_egrad, o_lap, o_enh, o_str,
              o_top, o_bot, x, w, h, e):
    """src/morsi.c:278-310: None outputs are skipped."""
    import ctypes
    import numpy as np
    outs = [o_ero, o_dil, o_ope, o_clo, o_grad, o_igrad, o_egrad, o_lap, o_enh, o_str, o_top, o_bot]
    xs = np.ascontiguousarray(x, dtype=np.float32)
    ptrs = (ctypes.c_void_p * 12)()
    for k, out in enumerate(outs):
        if out is not None:
            assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.size >= w * h
            ptrs[k] = out.ctypes.data
    ee = np.ascontiguousarray(e, dtype=np.int32)
    _b.check(_b.lib().morsi_cuda_apply_all(ee.ctypes.data_as(_b._i32p), xs.ctypes.data, ptrs, w, h, 1))
