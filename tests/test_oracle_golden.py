"""The CPU restatement against the golden dumps of the UNMODIFIED reference (run on a B200 by
tests/golden/make_golden.py).  This is what pins the oracle: the reference ships no tests or
vectors of its own for this path."""
import pytest

from tests import harness as H


@pytest.mark.parametrize("case", H.CASES, ids=lambda c: c.name)
def test_cpu_oracle_matches_reference_dump(case):
    want, digest = H.golden_as_calls(case)
    ref, query = case.inputs()
    assert H.inputs_digest(ref, query) == digest, "synthetic inputs are not reproducible from the seed"
    got = H.run_cpu_oracle(case, ref, query, max_hits_device=748058112)
    H.assert_calls_equal(got, want, "cpu-oracle vs reference golden")


def test_golden_covers_every_case_and_is_nontrivial():
    total_hsps = 0
    for case in H.CASES:
        assert H.golden_path(case).exists(), f"missing golden for {case.name}"
        calls, _ = H.load_golden(case)
        assert len(calls) >= 1
        total_hsps += sum(c[6].size for c in calls)
    assert total_hsps > 40000
    # the multi-iteration case really has more than two iterations' worth of hits
    calls, _ = H.load_golden(H.CASES_BY_NAME["diverged_multi_iter"])
    assert max(c[5] for c in calls) > 3 * H.CASES_BY_NAME["diverged_multi_iter"].max_hits_override
