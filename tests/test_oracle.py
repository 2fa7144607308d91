"""The oracle (oracle/morsi_oracle.c) against the golden vectors recorded from
the reference, and against the compiled reference itself where it exists."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import OPS, oracle, reference
from oracle.oracle import Reference
from tests.golden.make_golden import adversarial_input, kat_input, sha

HAVE_REF = os.path.exists(Reference.path)


def test_kat_input(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat_9_7.json")))
    x = kat_input()
    assert sha(x) == kat["input_sha"] == "bfffc36c57a1fc0e"
    assert x.astype(np.float64).sum() == kat["input_sum"]


def test_elements_match_reference_builders(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "elements.json")))
    o = oracle()
    for name, e in gold.items():
        got = o.element(name)
        if e is None:
            assert got is None, name
        else:
            assert got is not None and list(got) == e, name
    # geometry facts of SURVEY 9.2
    for name, n in [("cross", 5), ("square", 9), ("disk2", 9), ("disk3", 25), ("disk4.2", 57),
                    ("disk5", 69), ("disk7", 145), ("disk15", 697), ("dysk3", 16), ("dysk15", 88)]:
        assert o.element(name)[0] == n


def test_element_grammar_quirks():
    o = oracle()
    assert o.element("square3") is None          # src/morsi.c:498 exact strcmp
    assert o.element("disk1") is None            # radius must be > 1
    assert list(o.element("kids2.5")) == list(o.element("disk2.5"))   # strspn is a set match
    assert list(o.element("rrrr3")) == list(o.element("Drec3"))       # last test wins
    assert o.operation("tophat") == 13 and o.operation("top-hat") == -1


def test_known_answers(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "kat_9_7.json")))
    gold_e = json.load(open(os.path.join(golden_dir, "elements.json")))
    o, x = oracle(), kat_input()
    assert len(kat["cases"]) == 12 * 18
    for key, (h, s, y00, ymid) in kat["cases"].items():
        name, op = key.split()
        y = o.apply(op, np.array(gold_e[name], dtype=np.int32), x)
        assert sha(y) == h, key
        assert float(y.astype(np.float64).sum()) == s, key
    # the 14 vectors printed in SURVEY.md 9.7
    assert kat["cases"]["square erosion"][0] == "03b0352a05bb8b66"
    assert kat["cases"]["disk15 tophat"][0] == "b0d3afa47fee997e"
    assert kat["cases"]["disk5 median"][0] == "33f92592c1942319"


def _same_bits(a, b):
    """bit-exact except that NaNs compare by class (payloads are platform noise)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    an, bn = np.isnan(a), np.isnan(b)
    return np.array_equal(an, bn) and np.array_equal(a.view(np.uint32)[~an], b.view(np.uint32)[~bn])


def test_adversarial_bits(golden_dir):
    z = np.load(os.path.join(golden_dir, "adversarial.npz"))
    x = z["x"]
    assert _same_bits(x, adversarial_input())
    o = oracle()
    names = [k[2:] for k in z.files if k.startswith("e:")]
    assert len(names) == 12
    for name in names:
        e = z["e:" + name]
        gold = z["y:" + name].view(np.float32)
        for k, op in enumerate(OPS):
            assert _same_bits(o.apply(op, e, x), gold[k]), (name, op)


def test_median_even_rule_and_empty():
    o = oracle()
    x = np.arange(6, dtype=np.float32).reshape(2, 3)
    y = o.apply("median", o.element("disk7"), x)
    assert np.all(y == 3.5)                       # (a[3]+a[4])/2
    nan = np.full((3, 3), np.nan, dtype=np.float32)
    assert np.all(np.isnan(o.apply("median", o.element("cross"), nan)))
    assert np.all(o.apply("erosion", o.element("cross"), nan) == np.inf)
    assert np.all(o.apply("dilation", o.element("cross"), nan) == -np.inf)


def test_signed_zero_last_wins():
    o = oracle()
    row = np.array([[0.0, -0.0, 0.0, -0.0, -0.0, 0.0, 0.0, -0.0]], dtype=np.float32)
    want = [1, 0, 1, 1, 0, 0, 1, 1]               # SURVEY 9.1-Z probe
    for op in ("erosion", "dilation"):
        y = o.apply(op, o.element("hrec2"), row)
        assert list(np.signbit(y[0]).astype(int)) == want


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_vs_reference_random(seed):
    """Differential check against the compiled reference on fresh inputs."""
    rng = np.random.default_rng(seed)
    o, r = oracle(), reference()
    h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    x = adversarial_input(seed=100 + seed, h=h, w=w)
    if seed == 2:
        x = rng.random((h, w), dtype=np.float32)
    for name in ["cross", "square", "disk3.3", "disk5", "dysk4", "hrec6", "vrec3", "drec4", "Drec3"]:
        e = o.element(name)
        for op in OPS:
            assert _same_bits(o.apply(op, e, x), r.apply(op, e, x)), (name, op, h, w)
