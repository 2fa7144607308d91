"""The comparator networks of the shared-window median kernel (k_median.cu) are
generated tables: check the generator's own self-test (every network against a
brute-force median on random windows with ties) and that the committed header
is what the generator emits."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_median_nets", os.path.join(ROOT, "tools", "gen_median_nets.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_networks_select_the_median():
    assert _gen().self_test(trials=60)


def test_header_is_current():
    want = _gen().emit()
    have = open(os.path.join(ROOT, "imscript_b200", "csrc", "median_nets.cuh")).read()
    assert have == want, "run python tools/gen_median_nets.py"


def test_shapes_match_shapes_cuh():
    """SHAPES in the generator must equal the MORSI_SHAPE table of shapes.cuh"""
    import re
    txt = open(os.path.join(ROOT, "imscript_b200", "csrc", "shapes.cuh")).read()
    table = {}
    for m in re.finditer(r"MORSI_SHAPE\((\d+),\s*(\d+),\s*([0-9,\s]+)\)", txt):
        table[int(m.group(1))] = (int(m.group(2)), [int(v) for v in m.group(3).split(",")])
    for sid, (r, hw) in _gen().SHAPES.items():
        assert table[sid] == (r, hw), sid
