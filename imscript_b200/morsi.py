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
    table = [(o_ero, "erosion"), (o_dil, "dilation"), (o_ope, "opening"), (o_clo, "closing"),
             (o_grad, "gradient"), (o_igrad, "igradient"), (o_egrad, "egradient"),
             (o_lap, "laplacian"), (o_enh, "enhance"), (o_str, "oscillation"),
             (o_top, "tophat"), (o_bot, "bothat")]
    for out, op in table:
        if out is not None:
            _run(op, out, x, w, h, e)
