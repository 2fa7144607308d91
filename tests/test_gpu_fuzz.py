"""A short run of the randomised differential test (tests/fuzz_cases.py); the
long form (100 s, 4282 cases, 0 mismatches in round 1) is `python tests/fuzz_cases.py 100 7`."""
import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_against_oracle():
    from tests.fuzz_cases import run
    cases, bad = run(budget=20.0, seed=11, verbose=False)
    assert cases > 100
    assert not bad, bad[:5]
