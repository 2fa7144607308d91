#!/usr/bin/env python
"""Regenerate tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference/src/morsi.c).  Run in the dev container:

    python tests/golden/make_golden.py

Outputs (committed):
  kat_9_7.json        sha256[:16] / float64 sum / two probes for every
                      (element, operation) pair on the SURVEY 9.7 input;
  adversarial.npz     a 19x23 image drawn from {+0,-0,NaN,+-Inf,small ints,...}
                      plus the reference output bits of all 18 operations for
                      ten structuring elements (incl. raw user lists);
  elements.json       the offset lists the reference builders return;
  cli.json            exit codes / stderr / output-NPY hashes of the reference CLI.
The reference ships no vectors of its own for this path.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import OPS, oracle, reference  # noqa: E402

ELEMENT_NAMES = ["cross", "square", "disk2", "disk2.5", "disk3", "disk4.2", "disk5",
                 "disk7", "disk15", "dysk2", "dysk3", "dysk5", "hrec2", "hrec3",
                 "hrec7.5", "vrec2", "vrec4", "drec3", "Drec2.2", "Drec5"]
# raw user lists (the "user-defined mask" mechanism, src/morsi.c:48-54):
USER = {
    "user_offcentre": [4, 0, 1, -1, 0, 0, 2, 1, -1, 0, 3, -2],
    "user_repeats": [6, 0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 2, 0, 2, -2, -1],
    "user_single": [1, 0, 0, 0, 2, 1],
}


def kat_input():
    j, i = np.mgrid[0:48, 0:64]
    return (((i * 73 + j * 151 + (i * j) % 17) % 256).astype(np.float32)
            / np.float32(7)).astype(np.float32)


def adversarial_input(seed=7, h=19, w=23):
    rng = np.random.default_rng(seed)
    pool = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1, -1, 2, -2, 3, 0.5,
                     1e-45, -1e-45, 3.4e38, 0.0, -0.0, 1, 2], dtype=np.float32)
    return pool[rng.integers(0, pool.size, size=(h, w))]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    ref, ora = reference(), oracle()
    # -- elements
    elements = {}
    for name in ELEMENT_NAMES:
        e = ora.element(name)
        if name not in ("cross", "square"):
            er = ref.build(name[:4], float(name[4:]))
            assert np.array_equal(e, er), name
        elements[name] = [int(v) for v in e]
    for name in ["disk1", "disk0.5", "hrec1", "dysk1"]:
        assert ref.build(name[:4], float(name[4:])) is None
        elements[name] = None
    with open(os.path.join(HERE, "elements.json"), "w") as f:
        json.dump(elements, f, separators=(",", ":"))

    # -- known answers on the SURVEY 9.7 input
    x = kat_input()
    kat = {"input_sha": sha(x), "input_sum": float(x.astype(np.float64).sum()), "cases": {}}
    for name in ["cross", "square", "disk2.5", "disk4.2", "disk5", "disk7", "disk15",
                 "dysk3", "hrec3", "vrec4", "drec3", "Drec2.2"]:
        e = np.array(elements[name], dtype=np.int32)
        for op in OPS:
            y = ref.apply(op, e, x)
            kat["cases"][f"{name} {op}"] = [sha(y), float(y.astype(np.float64).sum()),
                                            float(y[0, 0]), float(y[24, 32])]
    with open(os.path.join(HERE, "kat_9_7.json"), "w") as f:
        json.dump(kat, f, indent=0)

    # -- adversarial values, full output bits
    xa = adversarial_input()
    arrays = {"x": xa}
    adv_elements = {k: elements[k] for k in ["cross", "square", "disk2.5", "disk3", "dysk3",
                                             "hrec3", "vrec2", "drec3", "Drec2.2"]}
    adv_elements.update(USER)
    for name, e in adv_elements.items():
        e = np.array(e, dtype=np.int32)
        arrays["e:" + name] = e
        out = np.stack([ref.apply(op, e, xa) for op in OPS])
        arrays["y:" + name] = out.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "adversarial.npz"), **arrays)

    # -- CLI behaviour
    cli = {}
    with tempfile.TemporaryDirectory() as d:
        fin = os.path.join(d, "in.npy")
        rgb = np.stack([kat_input(), kat_input()[::-1], kat_input()[:, ::-1]], axis=-1)
        np.save(fin, rgb)
        for argv in [["cross", "gradient", fin], ["disk3", "opening", fin],
                     ["square3", "erosion", fin], ["disk1", "erosion", fin],
                     ["square", "nosuchop", fin], ["kids2.5", "rank", fin],
                     ["rrrr3", "dilation", fin], [], ["a", "b", "c", "d", "e"],
                     ["--version"], ["--help"], ["-h"], ["-?"], ["--help-oneliner"]]:
            p = subprocess.run([ref.cli] + argv, capture_output=True)
            key = " ".join(a if a != fin else "IN" for a in argv)
            cli[key] = {"rc": p.returncode,
                        "stderr": p.stderr.decode().replace(ref.cli, "morsi"),
                        "stdout_sha": hashlib.sha256(p.stdout).hexdigest()[:16],
                        "stdout_len": len(p.stdout)}
            if argv and argv[0].startswith("-"):
                cli[key]["stdout"] = p.stdout.decode()
    with open(os.path.join(HERE, "cli.json"), "w") as f:
        json.dump(cli, f, indent=0)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
