#!/usr/bin/env python
"""Generates imscript_b200/csrc/median_nets.cuh: the compile-time tables of the
shared-window median kernel (k_median.cu).

A thread computes the medians of a 2x2 block of pixels.  For a row-run element
D (shapes.cuh) with N = 2h+1 samples, the four windows share a CORE (samples in
all four), each horizontal pair of windows shares EXT more, and every window
has UNI samples of its own.  Because the median of N values has h values on
either side, the DROP = |CORE| - (h+1) smallest and largest CORE values cannot
be the median of any of the four windows ("forgetful selection"): the CORE is
sorted once (Batcher's odd-even merge sort, only the KEEP middle outputs are
live, the compiler removes the rest), and each pixel merges its sorted EXT and
UNI samples and selects rank |EXT|+|UNI| of KEEP + EXT + UNI values with
min_i max(A[i], C[n-1-i]).  FULL is a sorting network for all N samples: windows
with missing or non-finite samples (image border, NaN/Inf data) sort their
samples with the absent ones replaced by +INF and pick the reference's
variable-count positions (src/morsi.c:91-101) from the sorted order.

    python tools/gen_median_nets.py          # rewrites the header
    python tools/gen_median_nets.py --check  # self-test of the networks only
"""
import os
import random
import sys

SHAPES = {  # id: (reach, half-widths per row) -- must match shapes.cuh
    0: (2, [1, 2, 2, 2, 1]),
    1: (2, [2, 2, 2, 2, 2]),
    2: (3, [1, 2, 3, 3, 3, 2, 1]),
    3: (3, [2, 3, 3, 3, 3, 3, 2]),
    4: (4, [1, 2, 3, 4, 4, 4, 3, 2, 1]),
    5: (4, [2, 3, 4, 4, 4, 4, 4, 3, 2]),
    6: (5, [1, 3, 4, 4, 5, 5, 5, 4, 4, 3, 1]),
    7: (5, [3, 4, 5, 5, 5, 5, 5, 5, 5, 4, 3]),
    8: (6, [3, 4, 5, 6, 6, 6, 6, 6, 6, 6, 5, 4, 3]),
}


def batcher_sort(n):
    """Batcher's odd-even merge sort for any n (iterative form)."""
    out, p = [], 1
    while p < n:
        k = p
        while k >= 1:
            j = k % p
            while j + k < n:
                for i in range(min(k, n - j - k)):
                    if (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        out.append((i + j, i + j + k))
                j += 2 * k
            k //= 2
        p *= 2
    return out


def odd_even_merge(ia, ib, comps):
    """Merges the sorted sequences held at positions ia and ib; returns the
    positions in sorted order (Knuth 5.3.4, any lengths)."""
    if not ia:
        return list(ib)
    if not ib:
        return list(ia)
    if len(ia) == 1 and len(ib) == 1:
        comps.append((ia[0], ib[0]))
        return [ia[0], ib[0]]
    c = odd_even_merge(ia[0::2], ib[0::2], comps)
    d = odd_even_merge(ia[1::2], ib[1::2], comps)
    z, i = [c[0]], 1
    while i < len(c) and i - 1 < len(d):
        comps.append((d[i - 1], c[i]))
        z += [d[i - 1], c[i]]
        i += 1
    return z + c[i:] + d[i - 1:]


def plan(R, hw):
    D = [(dx, dy) for dy in range(-R, R + 1) for dx in range(-hw[dy + R], hw[dy + R] + 1)]
    N = len(D)
    h = N // 2
    W = {(a, b): {(dx + a, dy + b) for dx, dy in D} for a in (0, 1) for b in (0, 1)}
    core = sorted(W[0, 0] & W[1, 0] & W[0, 1] & W[1, 1], key=lambda t: (t[1], t[0]))
    pair = [W[0, b] & W[1, b] for b in (0, 1)]
    ext = [sorted(pair[b] - set(core), key=lambda t: (t[1], t[0])) for b in (0, 1)]
    uni = [sorted(W[a, b] - pair[b], key=lambda t: (t[1], t[0])) for b in (0, 1) for a in (0, 1)]  # index 2b+a
    drop = len(core) - (h + 1)
    assert N % 2 == 1 and drop >= 0
    keep = len(core) - 2 * drop
    ne, nu = len(ext[0]), len(uni[0])
    assert len(ext[1]) == ne and all(len(u) == nu for u in uni) and keep == ne + nu + 1
    mcomps = []
    order = odd_even_merge(list(range(ne)), list(range(ne, ne + nu)), mcomps)
    return dict(N=N, h=h, core=core, ext=ext, uni=uni, drop=drop, keep=keep,
                core_net=batcher_sort(len(core)), ext_net=batcher_sort(ne), uni_net=batcher_sort(nu),
                full_net=batcher_sort(N),
                mrg_net=mcomps, mrg_order=order, D=D)


def run_net(v, net):
    for a, b in net:
        if v[a] > v[b]:
            v[a], v[b] = v[b], v[a]


def self_test(trials=300):
    rng = random.Random(1)
    for sid, (R, hw) in SHAPES.items():
        P = plan(R, hw)
        for _ in range(trials):
            img = {}
            span = range(-R - 1, R + 3)
            for y in span:
                for x in span:
                    img[x, y] = rng.choice([rng.random(), float(rng.randrange(4))])
            cv = [img[t] for t in P["core"]]
            run_net(cv, P["core_net"])
            A = cv[P["drop"]:P["drop"] + P["keep"]]
            for b in (0, 1):
                ev = [img[t] for t in P["ext"][b]]
                run_net(ev, P["ext_net"])
                for a in (0, 1):
                    uv = [img[t] for t in P["uni"][2 * b + a]]
                    run_net(uv, P["uni_net"])
                    m = ev + uv
                    run_net(m, P["mrg_net"])
                    C = [m[o] for o in P["mrg_order"]]
                    n = len(C)
                    got = A[n]
                    for t in range(n):
                        got = min(got, max(A[t], C[n - 1 - t]))
                    want = sorted(img[dx + a, dy + b] for dx, dy in P["D"])[P["h"]]
                    assert got == want, (sid, a, b, got, want)
    return True


def c_array(name, typ, vals, dims):
    # a constexpr function around a local table: usable in device code, where a
    # static constexpr member array is not
    flat = ", ".join(str(v) for v in vals)
    return (f"\t__host__ __device__ static constexpr int {name}(int i) {{ constexpr {typ} t{dims} = {{{flat}}}; "
            f"return t[i]; }}\n")


def emit():
    out = ["// median_nets.cuh -- GENERATED by tools/gen_median_nets.py; do not edit.\n",
           "// Tables of the shared-window median kernel (k_median.cu): sample offsets\n",
           "// relative to the top-left pixel of a 2x2 block and comparator networks.\n",
           "#pragma once\n\ntemplate <int ID> struct MedNet { static constexpr bool ok = false; };\n"]
    for sid, (R, hw) in SHAPES.items():
        P = plan(R, hw)
        nc, ne, nu = len(P["core"]), len(P["ext"][0]), len(P["uni"][0])
        out.append(f"\ntemplate <> struct MedNet<{sid}> {{\n\tstatic constexpr bool ok = true;\n")
        out.append(f"\tstatic constexpr int N = {P['N']}, NCORE = {nc}, NEXT = {ne}, NUNI = {nu}, DROP = {P['drop']}, KEEP = {P['keep']};\n")
        out.append(f"\tstatic constexpr int NCORE_NET = {len(P['core_net'])}, NEXT_NET = {len(P['ext_net'])}, "
                   f"NUNI_NET = {len(P['uni_net'])}, NMRG_NET = {len(P['mrg_net'])}, NFULL_NET = {len(P['full_net'])};\n")
        out.append(c_array("core_dx", "signed char", [t[0] for t in P["core"]], f"[{nc}]"))
        out.append(c_array("core_dy", "signed char", [t[1] for t in P["core"]], f"[{nc}]"))
        out.append(c_array("ext_dx", "signed char", [t[0] for b in (0, 1) for t in P["ext"][b]], f"[{2 * ne}]"))
        out.append(c_array("ext_dy", "signed char", [t[1] for b in (0, 1) for t in P["ext"][b]], f"[{2 * ne}]"))
        out.append(c_array("uni_dx", "signed char", [t[0] for q in range(4) for t in P["uni"][q]], f"[{4 * nu}]"))
        out.append(c_array("uni_dy", "signed char", [t[1] for q in range(4) for t in P["uni"][q]], f"[{4 * nu}]"))
        for nm in ("core_net", "ext_net", "uni_net", "mrg_net", "full_net"):
            net = P[nm] or [(0, 0)]
            out.append(c_array(nm + "_a", "unsigned char", [c[0] for c in net], f"[{len(net)}]"))
            out.append(c_array(nm + "_b", "unsigned char", [c[1] for c in net], f"[{len(net)}]"))
        out.append(c_array("mrg_order", "unsigned char", P["mrg_order"], f"[{ne + nu}]"))
        out.append("};\n")
    return "".join(out)


if __name__ == "__main__":
    assert self_test()
    if "--check" in sys.argv:
        print("median networks ok")
    else:
        dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "imscript_b200", "csrc", "median_nets.cuh")
        open(dst, "w").write(emit())
        print("wrote", os.path.normpath(dst))
