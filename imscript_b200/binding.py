"""ctypes binding of include/morsi_cuda.h."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = ["erosion", "dilation", "median", "rank", "opening", "closing",
       "gradient", "igradient", "egradient", "laplacian", "enhance", "blur",
       "oscillation", "tophat", "bothat", "iblur", "eblur", "cblur"]   # src/morsi.c:510-527

# every symbol include/morsi_cuda.h declares for libmorsi_cuda.so
EXPORTED_SYMBOLS = [
    "morsi_cuda_strerror", "morsi_cuda_last_error", "morsi_element_parse",
    "morsi_build_disk", "morsi_build_dysk", "morsi_build_hrec", "morsi_build_vrec",
    "morsi_build_drec", "morsi_build_Drec", "morsi_element_free", "morsi_operation_parse",
    "morsi_operation_name", "morsi_element_describe", "morsi_cuda_device_count",
    "morsi_cuda_init", "morsi_cuda_shutdown", "morsi_cuda_apply", "morsi_cuda_apply_all", "morsi_cuda_apply_device",
    "morsi_cuda_apply_band_device", "morsi_cuda_halo_rows", "morsi_cuda_set_path",
    "morsi_cuda_launch_count", "morsi_cuda_launch_count_reset", "morsi_cuda_malloc",
    "morsi_cuda_free", "morsi_cuda_host_alloc", "morsi_cuda_host_free", "morsi_cuda_memcpy_h2d",
    "morsi_cuda_memcpy_d2h", "morsi_cuda_memcpy_d2d", "morsi_cuda_sync", "morsi_cuda_synth",
    "morsi_synth_host", "morsi_cuda_event_create", "morsi_cuda_event_record",
    "morsi_cuda_event_elapsed_ms", "morsi_cuda_event_destroy",
    "morsi_shard_create", "morsi_shard_handle", "morsi_shard_connect", "morsi_shard_rows",
    "morsi_shard_buffer", "morsi_shard_stream", "morsi_shard_apply", "morsi_shard_apply_host",
    "morsi_cuda_apply_all_device", "morsi_cuda_apply_interleaved", "morsi_cuda_apply_stream", "morsi_cuda_apply_chain",
    "morsi_shard_exchange", "morsi_cuda_stream_create", "morsi_cuda_stream_destroy",
    "morsi_shard_sync", "morsi_shard_halo_bytes", "morsi_shard_destroy", "morsi_cuda_apply_sharded",
]
SHARD_HANDLE_BYTES = 128
COMPAT_SYMBOLS = ["morsi_" + o for o in OPS] + ["morsi_all", "build_disk"]

READ_ROWS_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p)
WRITE_ROWS_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p


class MorsiError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"libmorsi_cuda error {code}: {what}")
        self.code = code


def lib_path(name="libmorsi_cuda.so"):
    # MORSI_CUDA_LIB: an alternative build of the library (A/B measurements of kernel variants)
    if name == "libmorsi_cuda.so" and os.environ.get("MORSI_CUDA_LIB"):
        return os.environ["MORSI_CUDA_LIB"]
    return os.path.join(HERE, "lib", name)


_lib = None


def lib():
    """The loaded shared library; raises if it has not been built
    (python -c 'import __graft_entry__ as g; g.build()' or `make`)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise MorsiError(-1, f"{path} is missing: build it with `make` (nvcc, sm_100a); "
                                 "there is no CPU fallback")
        L = ctypes.CDLL(path)
        L.morsi_cuda_strerror.restype = ctypes.c_char_p
        L.morsi_cuda_last_error.restype = ctypes.c_char_p
        L.morsi_operation_name.restype = ctypes.c_char_p
        L.morsi_operation_name.argtypes = [ctypes.c_int]
        L.morsi_cuda_launch_count.restype = ctypes.c_long
        L.morsi_element_parse.argtypes = [ctypes.c_char_p, ctypes.POINTER(_i32p)]
        for k in ("disk", "dysk", "hrec", "vrec", "drec", "Drec"):
            f = getattr(L, "morsi_build_" + k)
            f.restype = _i32p
            f.argtypes = [ctypes.c_float]
        L.morsi_element_free.argtypes = [_i32p]
        L.morsi_operation_parse.argtypes = [ctypes.c_char_p]
        L.morsi_element_describe.argtypes = [_i32p, ctypes.c_char_p, ctypes.c_size_t]
        L.morsi_cuda_apply.argtypes = [ctypes.c_int, _i32p, _vp, _vp,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.morsi_cuda_apply_all.argtypes = [_i32p, _vp, ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.morsi_cuda_apply_device.argtypes = [ctypes.c_int, _i32p, _vp, _vp, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, _vp]
        L.morsi_cuda_apply_band_device.argtypes = [ctypes.c_int, _i32p, _vp, ctypes.c_int, ctypes.c_int,
                                                   _vp, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, _vp]
        L.morsi_cuda_halo_rows.argtypes = [ctypes.c_int, _i32p, _i32p, _i32p]
        L.morsi_cuda_malloc.argtypes = [ctypes.POINTER(_vp), ctypes.c_size_t]
        L.morsi_cuda_free.argtypes = [_vp]
        L.morsi_cuda_host_alloc.argtypes = [ctypes.POINTER(_vp), ctypes.c_size_t]
        L.morsi_cuda_host_free.argtypes = [_vp]
        for k in ("h2d", "d2h", "d2d"):
            getattr(L, "morsi_cuda_memcpy_" + k).argtypes = [_vp, _vp, ctypes.c_size_t, _vp]
        L.morsi_cuda_sync.argtypes = [_vp]
        L.morsi_cuda_synth.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_uint, ctypes.c_int, _vp]
        L.morsi_synth_host.restype = None
        L.morsi_synth_host.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_uint, ctypes.c_int]
        L.morsi_cuda_event_create.argtypes = [ctypes.POINTER(_vp)]
        L.morsi_cuda_event_record.argtypes = [_vp, _vp]
        L.morsi_cuda_event_elapsed_ms.argtypes = [_vp, _vp, ctypes.POINTER(ctypes.c_float)]
        L.morsi_cuda_event_destroy.argtypes = [_vp]
        L.morsi_cuda_apply_all_device.argtypes = [_i32p, _vp, ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]
        L.morsi_cuda_apply_interleaved.argtypes = [ctypes.c_int, _i32p, _vp, _vp, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int]
        L.morsi_cuda_apply_stream.argtypes = [ctypes.c_int, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              READ_ROWS_FN, WRITE_ROWS_FN, _vp]
        L.morsi_shard_create.argtypes = [ctypes.POINTER(_vp)] + [ctypes.c_int] * 7
        L.morsi_shard_handle.argtypes = [_vp, _vp]
        L.morsi_shard_connect.argtypes = [_vp, _vp]
        L.morsi_shard_rows.argtypes = [_vp, _i32p, _i32p, _i32p, _i32p]
        L.morsi_shard_buffer.restype = _vp
        L.morsi_shard_buffer.argtypes = [_vp, ctypes.c_int]
        L.morsi_shard_stream.restype = _vp
        L.morsi_shard_stream.argtypes = [_vp]
        L.morsi_shard_apply.argtypes = [_vp, ctypes.c_int, _i32p, ctypes.c_int, ctypes.c_int]
        L.morsi_shard_apply_host.argtypes = [_vp, ctypes.c_int, _i32p, _vp, _vp]
        L.morsi_shard_sync.argtypes = [_vp]
        L.morsi_shard_exchange.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.morsi_cuda_stream_create.argtypes = [ctypes.POINTER(_vp)]
        L.morsi_cuda_stream_destroy.argtypes = [_vp]
        L.morsi_shard_halo_bytes.restype = ctypes.c_longlong
        L.morsi_shard_halo_bytes.argtypes = [_vp]
        L.morsi_shard_destroy.argtypes = [_vp]
        L.morsi_cuda_apply_sharded.argtypes = [ctypes.c_int, _i32p, _vp, _vp, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        L = lib()
        raise MorsiError(rc, f"{L.morsi_cuda_strerror(rc).decode()}: {L.morsi_cuda_last_error().decode()}")


def _op(op):
    return OPS.index(op) if isinstance(op, str) else int(op)


def _e(e):
    if isinstance(e, str):
        e = parse_element(e)
        if e is None:
            raise MorsiError(1, "elements = cross, square ...")
    e = np.ascontiguousarray(e, dtype=np.int32)
    if e.ndim != 1 or e.size < 4 or e.size < 4 + 2 * int(e[0]):
        raise MorsiError(1, "malformed structuring element list")
    return e


def _take(ptr):
    if not ptr:
        return None
    n = 4 + 2 * ptr[0]
    out = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
    lib().morsi_element_free(ptr)
    return out


def parse_element(name):
    """src/morsi.c:496-508 -> int32 list, or None where the reference errors out."""
    p = _i32p()
    rc = lib().morsi_element_parse(name.encode(), ctypes.byref(p))
    return _take(p) if rc == 0 else None


def build_element(kind, radius):
    return _take(getattr(lib(), "morsi_build_" + kind)(radius))


def parse_operation(name):
    return lib().morsi_operation_parse(name.encode())


def describe_element(e):
    e = _e(e)
    buf = ctypes.create_string_buffer(128)
    check(lib().morsi_element_describe(e.ctypes.data_as(_i32p), buf, 128))
    return buf.value.decode()


def device_count():
    return lib().morsi_cuda_device_count()


def halo_rows(op, e):
    e = _e(e)
    up, down = ctypes.c_int(), ctypes.c_int()
    check(lib().morsi_cuda_halo_rows(_op(op), e.ctypes.data_as(_i32p), ctypes.byref(up), ctypes.byref(down)))
    return up.value, down.value


def apply(op, e, x):
    """Host arrays in, host array out: morsi_cuda_apply().  x is (h,w) or
    (planes,h,w) float32 (planar, as iio_read_image_float_split returns it)."""
    e = _e(e)
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim < 2:
        raise MorsiError(1, "image must have at least 2 dimensions")
    h, w = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    y = np.empty_like(x)
    check(lib().morsi_cuda_apply(_op(op), e.ctypes.data_as(_i32p), x.ctypes.data, y.ctypes.data, w, h, planes))
    return y


class DeviceBuffer:
    """A device allocation owned through the C ABI (morsi_cuda_malloc)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = _vp()
        check(lib().morsi_cuda_malloc(ctypes.byref(p), self.nbytes))
        self.ptr = p.value

    @classmethod
    def from_host(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        check(lib().morsi_cuda_memcpy_h2d(b.ptr, a.ctypes.data, a.nbytes, None))
        check(lib().morsi_cuda_sync(None))
        return b

    def to_host(self, shape, dtype=np.float32):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().morsi_cuda_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, None))
        check(lib().morsi_cuda_sync(None))
        return out

    def free(self):
        if self.ptr:
            lib().morsi_cuda_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def apply_device(op, e, d_x, d_y, w, h, planes=1, stream=None, sync=True):
    e = _e(e)
    px = d_x.ptr if isinstance(d_x, DeviceBuffer) else d_x
    py = d_y.ptr if isinstance(d_y, DeviceBuffer) else d_y
    check(lib().morsi_cuda_apply_device(_op(op), e.ctypes.data_as(_i32p), px, py, w, h, planes, stream))
    if sync:
        check(lib().morsi_cuda_sync(stream))


def apply_band_device(op, e, d_x, x_row0, x_rows, d_y, y_row0, y_rows, w, h, stream=None, sync=True):
    e = _e(e)
    px = d_x.ptr if isinstance(d_x, DeviceBuffer) else d_x
    py = d_y.ptr if isinstance(d_y, DeviceBuffer) else d_y
    check(lib().morsi_cuda_apply_band_device(_op(op), e.ctypes.data_as(_i32p), px, x_row0, x_rows,
                                             py, y_row0, y_rows, w, h, stream))
    if sync:
        check(lib().morsi_cuda_sync(stream))


def apply_interleaved(op, e, x):
    """x: (h, w, pd) uint8 / uint16 / float32, pixel-interleaved -> (h, w, pd) float32."""
    e = _e(e)
    x = np.ascontiguousarray(x)
    if x.ndim == 2:
        x = x[:, :, None]
    t = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}[x.dtype]
    h, w, pd = x.shape
    y = np.empty((h, w, pd), np.float32)
    check(lib().morsi_cuda_apply_interleaved(_op(op), e.ctypes.data_as(_i32p), x.ctypes.data, y.ctypes.data, w, h, pd, t))
    return y


def apply_stream(op, e, w, h, planes, read_rows, write_rows):
    """read_rows(plane, row0, nrows) -> (nrows, w) float32 array; write_rows(plane, row0, rows_array)."""
    e = _e(e)

    def rd(user, plane, row0, nrows, dst):
        try:
            a = np.ascontiguousarray(read_rows(plane, row0, nrows), dtype=np.float32)
            ctypes.memmove(dst, a.ctypes.data, a.nbytes)
            return 0
        except Exception:
            return 1

    def wr(user, plane, row0, nrows, src):
        try:
            a = np.ctypeslib.as_array(ctypes.cast(src, _f32p), shape=(nrows, w)).copy()
            write_rows(plane, row0, a)
            return 0
        except Exception:
            return 1
    check(lib().morsi_cuda_apply_stream(_op(op), e.ctypes.data_as(_i32p), w, h, planes,
                                        READ_ROWS_FN(rd), WRITE_ROWS_FN(wr), None))


class Quantizer(ctypes.Structure):
    _fields_ = [("black", ctypes.c_float), ("white", ctypes.c_float), ("to_uint8", ctypes.c_int)]


def apply_chain(steps, x, quant=None):
    """steps: [(op, element), ...] applied in order on the device; quant: None or (black, white, to_uint8)
    = qeasy (src/qeasy.c) as the last stage.  x: (h,w) or (planes,h,w) float32 -> same shape, float32 or uint8."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    h, w = x.shape[-2:]
    planes = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    es = [_e(e) for _, e in steps]
    ops = (ctypes.c_int * len(steps))(*[_op(o) for o, _ in steps])
    eps = (_i32p * len(steps))(*[e.ctypes.data_as(_i32p) for e in es])
    q = None
    y = np.empty_like(x)
    if quant is not None:
        q = Quantizer(float(quant[0]), float(quant[1]), int(bool(quant[2])))
        if quant[2]:
            y = np.empty(x.shape, np.uint8)
    L = lib()
    L.morsi_cuda_apply_chain.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_i32p), _vp, _vp,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Quantizer)]
    check(L.morsi_cuda_apply_chain(len(steps), ops, eps, x.ctypes.data, y.ctypes.data, w, h, planes,
                                   ctypes.byref(q) if q is not None else None))
    return y


def apply_sharded(op, e, x, ndev, iterations=1):
    """One process, `ndev` devices: morsi_cuda_apply_sharded() on a (h,w) plane."""
    e = _e(e)
    x = np.ascontiguousarray(x, dtype=np.float32)
    h, w = x.shape
    y = np.empty_like(x)
    check(lib().morsi_cuda_apply_sharded(_op(op), e.ctypes.data_as(_i32p), x.ctypes.data, y.ctypes.data,
                                         w, h, ndev, iterations))
    return y


def synth_host(w, rows, row0=0, plane=0, seed=1, dist=0):
    out = np.empty((rows, w), dtype=np.float32)
    lib().morsi_synth_host(out.ctypes.data, w, rows, row0, plane, seed, dist)
    return out
