"""Randomised differential test of the C ABI against the oracle: random sizes
(incl. widths that are not multiples of 4), elements from every family, all 18
operations, value distributions 0/1/2, whole images and row bands.
Used by tests/test_gpu_fuzz.py (short) and `python tests/fuzz_cases.py [seconds] [seed]` (long)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imscript_b200 as M                      # noqa: E402
from oracle import OPS, oracle                 # noqa: E402

ELEMENTS = ["cross", "square", "disk2.5", "disk3", "disk3.5", "disk4", "disk4.2", "disk5", "disk5.1", "disk6", "disk7",
            "disk8", "disk9", "disk10", "disk11", "disk12", "disk13", "disk14", "disk15", "disk6.5", "disk16",
            "dysk3", "dysk5", "dysk8", "hrec2", "hrec7", "hrec33", "hrec60", "vrec2", "vrec9", "vrec30", "vrec55", "drec5", "Drec6"]


def same(a, b):
    an, bn = np.isnan(a), np.isnan(b)
    return np.array_equal(an, bn) and np.array_equal(a.view(np.uint32)[~an], b.view(np.uint32)[~bn])


def run(budget=60.0, seed=1, verbose=True):
    """returns (cases, list of mismatch descriptions)"""
    o = oracle()
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n, bad = 0, []
    while time.time() - t0 < budget:
        name = ELEMENTS[rng.integers(len(ELEMENTS))]
        e = o.element(name)
        big = e[0] > 200
        w = int(rng.integers(1, 700 if not big else 420))
        h = int(rng.integers(1, 300 if not big else 200))
        if rng.random() < 0.3:
            w = (w + 3) // 4 * 4
        planes = int(rng.integers(1, 4))
        dist = int(rng.integers(0, 3))
        op = OPS[rng.integers(len(OPS))]
        if op in ("median", "rank") and e[0] > 150 and w * h > 20000:
            continue
        x = np.stack([M.synth_host(w, h, plane=p, seed=int(rng.integers(1 << 20)), dist=dist) for p in range(planes)])
        if rng.random() < 0.5:
            x[x == 0] = 0.0          # without -0.0 the fast kernels must be right on their own
        want = o.apply(op, e, x)
        got = M.apply(op, e, x)
        n += 1
        if not same(got, want):
            bad.append(f"{name} {op} {w}x{h}x{planes} dist={dist}: "
                       f"{int((got.view(np.uint32) != want.view(np.uint32)).sum())} samples differ")
        # a random row band of plane 0 through the band entry point
        if h > 8 and rng.random() < 0.5:
            up, down = M.halo_rows(op, e)
            b0 = int(rng.integers(0, h - 1))
            b1 = int(rng.integers(b0 + 1, h + 1))
            i0, i1 = max(0, b0 - up), min(h, b1 + down)
            dx = M.DeviceBuffer.from_host(np.ascontiguousarray(x[0, i0:i1]))
            dy = M.DeviceBuffer((b1 - b0) * w * 4)
            M.apply_band_device(op, e, dx, i0, i1 - i0, dy, b0, b1 - b0, w, h)
            gb = dy.to_host((b1 - b0, w))
            n += 1
            if not same(gb, want[0, b0:b1]):
                bad.append(f"band [{b0},{b1}) {name} {op} {w}x{h} dist={dist}")
    if verbose:
        for b in bad:
            print("MISMATCH", b)
        print(f"fuzz: {n} cases, {len(bad)} mismatches, {time.time() - t0:.0f} s")
    return n, bad


if __name__ == "__main__":
    cases, mism = run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    sys.exit(1 if mism else 0)
