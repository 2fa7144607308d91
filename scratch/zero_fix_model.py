"""Design check for the planned signed-zero fix-up pass:
a value-based minimum (PTX min.f32: -0 < +0, NaN ignored) followed by a pass that
re-decides the sign of the outputs that are +-0 by scanning their window in
element order must equal the reference's erosion (last occurrence wins).
numpy model against the oracle on adversarial images; CPU only.
    python scratch/zero_fix_model.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle

o = oracle()
rng = np.random.default_rng(3)
vals = np.array([0.0, -0.0, 1.0, -1.0, 0.5, np.nan, np.inf, -np.inf, 2.0, -0.0, 0.0], dtype=np.float32)


def fast_then_fix(x, e, ismax):
    h, w = x.shape
    n = e[0]
    offs = [(int(e[4 + 2 * k] - e[2]), int(e[5 + 2 * k] - e[3])) for k in range(n)]
    y = np.empty_like(x)
    for j in range(h):
        for i in range(w):
            win = [x[j + dy, i + dx] for dx, dy in offs if 0 <= i + dx < w and 0 <= j + dy < h]
            win = [v for v in win if not np.isnan(v)]
            if not win:
                y[j, i] = -np.inf if ismax else np.inf
                continue
            # value-based extremum with -0 < +0
            key = lambda v: (v, 0 if np.signbit(v) else 1) if v == 0 else (v, 0)
            r = max(win, key=key) if ismax else min(win, key=key)
            if r == 0:                              # the fix-up: the last zero in element order carries the sign
                r = [v for v in win if v == 0][-1]
            y[j, i] = r
    return y


bad = 0
for name in ("cross", "square", "disk3", "dysk3", "hrec3", "drec3"):
    e = o.element(name)
    for trial in range(6):
        x = vals[rng.integers(len(vals), size=(9, 11))]
        for op, ismax in (("erosion", False), ("dilation", True)):
            want = o.apply(op, e, x)
            got = fast_then_fix(x, e, ismax)
            same = np.array_equal(np.isnan(got), np.isnan(want)) and \
                np.array_equal(got.view(np.uint32)[~np.isnan(got)], want.view(np.uint32)[~np.isnan(want)])
            bad += not same
print("zero-sign fix-up model:", "agrees with the reference on every case" if not bad else f"{bad} mismatches")
sys.exit(1 if bad else 0)
