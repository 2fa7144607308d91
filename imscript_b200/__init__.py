"""imscript_b200 -- B200 (sm_100a) implementation of imscript's `morsi` hot path.

This package is only the Python-side binding of the C ABI declared in
include/morsi_cuda.h (ctypes, no torch types): the product is
imscript_b200/lib/libmorsi_cuda.so plus the `morsi` host program.  There is no
CPU fallback: importing works without a GPU (so that the build can be checked),
every compute call raises MorsiError without one.
"""
from .binding import (OPS, MorsiError, DeviceBuffer, apply, apply_device, apply_band_device, apply_sharded, apply_interleaved, apply_stream, apply_chain,  # noqa: F401
                      build_element, describe_element, device_count, halo_rows, lib, lib_path,
                      parse_element, parse_operation, synth_host, EXPORTED_SYMBOLS)
from . import morsi  # noqa: F401
