"""ctypes front-ends for the two CPU checkers.

* ``Oracle``    -- oracle/_build/libmorsi_oracle.so, the restatement
                   (oracle/morsi_oracle.c), always buildable.
* ``Reference`` -- oracle/_ref/libmorsi_ref.so, the UNMODIFIED reference
                   functions of src/morsi.c compiled from /root/reference
                   (built in the dev container, shipped prebuilt to the GPU box).

TEST INFRASTRUCTURE ONLY: never imported by imscript_b200/.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = ["erosion", "dilation", "median", "rank", "opening", "closing",
       "gradient", "igradient", "egradient", "laplacian", "enhance", "blur",
       "oscillation", "tophat", "bothat", "iblur", "eblur", "cblur"]
KINDS = ["disk", "dysk", "hrec", "vrec", "drec", "Drec"]

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build_all(quiet=True):
    """(Re)build the checker libraries; compiling the checker is not using it."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _as_e(e):
    e = np.ascontiguousarray(e, dtype=np.int32)
    assert e.ndim == 1 and e.size >= 4 and e.size >= 4 + 2 * int(e[0])
    return e


class _Lib:
    prefix = None
    path = None

    def __init__(self):
        if not os.path.exists(self.path):
            build_all()
        self.lib = ctypes.CDLL(self.path)
        ap = getattr(self.lib, self.prefix + "_apply")
        ap.restype = ctypes.c_int
        ap.argtypes = [ctypes.c_int, _i32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int]
        self._apply = ap
        bd = getattr(self.lib, self.prefix + "_build")
        bd.restype = ctypes.c_int
        bd.argtypes = [ctypes.c_int, ctypes.c_float, _i32p, ctypes.c_int]
        self._build = bd

    def apply(self, op, e, x):
        """x: (h,w) or (planes,h,w) float32 -> same shape, one op per plane
        (the channel loop of src/morsi.c:539-543)."""
        if isinstance(op, str):
            op = OPS.index(op)
        e = _as_e(e)
        x = np.ascontiguousarray(x, dtype=np.float32)
        planes = x.reshape((-1,) + x.shape[-2:])
        y = np.empty_like(planes)
        h, w = planes.shape[-2:]
        for k in range(planes.shape[0]):
            rc = self._apply(op, e.ctypes.data_as(_i32p),
                             planes[k].ctypes.data_as(_f32p),
                             y[k].ctypes.data_as(_f32p), w, h)
            assert rc == 0
        return y.reshape(x.shape)

    def build(self, kind, radius):
        """-> int32 element list, or None where the reference returns NULL."""
        if isinstance(kind, str):
            kind = KINDS.index(kind)
        side = int(2 * abs(radius) + 8)
        cap = 2 * side * side + 8
        out = np.zeros(cap, dtype=np.int32)
        n = self._build(kind, radius, out.ctypes.data_as(_i32p), cap)
        assert n >= 0
        return out[:n].copy() if n else None


class Oracle(_Lib):
    prefix = "morsi_oracle"
    path = os.path.join(HERE, "_build", "libmorsi_oracle.so")

    def __init__(self):
        super().__init__()
        pe = self.lib.morsi_oracle_parse_element
        pe.restype = ctypes.c_int
        pe.argtypes = [ctypes.c_char_p, _i32p, ctypes.c_int]
        po = self.lib.morsi_oracle_parse_operation
        po.restype = ctypes.c_int
        po.argtypes = [ctypes.c_char_p]

    def element(self, name):
        """Element-name grammar of src/morsi.c:496-508 -> list or None."""
        try:
            r = float(name[4:]) if len(name) > 4 else 0.0
        except ValueError:
            r = 0.0
        side = int(2 * abs(r) + 8)
        cap = max(64, 2 * side * side + 8)
        out = np.zeros(cap, dtype=np.int32)
        n = self.lib.morsi_oracle_parse_element(name.encode(), out.ctypes.data_as(_i32p), cap)
        assert n >= 0
        return out[:n].copy() if n else None

    def operation(self, name):
        return self.lib.morsi_oracle_parse_operation(name.encode())


class Reference(_Lib):
    prefix = "morsi_ref"
    path = os.path.join(HERE, "_ref", "libmorsi_ref.so")
    cli = os.path.join(HERE, "_ref", "morsi_ref")

    def __init__(self):
        super().__init__()
        tm = self.lib.morsi_ref_time
        tm.restype = ctypes.c_double
        tm.argtypes = [ctypes.c_int, _i32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int]

    def time(self, op, e, x):
        """Seconds inside the reference function for one (h,w) plane."""
        if isinstance(op, str):
            op = OPS.index(op)
        e = _as_e(e)
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        h, w = x.shape
        return self.lib.morsi_ref_time(op, e.ctypes.data_as(_i32p),
                                       x.ctypes.data_as(_f32p),
                                       y.ctypes.data_as(_f32p), w, h)


def have_reference():
    return os.path.exists(Reference.path) or os.path.exists("/root/reference/src/morsi.c")


_o = _r = None


def oracle():
    global _o
    if _o is None:
        _o = Oracle()
    return _o


def reference():
    global _r
    if _r is None:
        _r = Reference()
    return _r
