"""
The oracle against the reference's own golden vectors.
In 320-bit mode the oracle must reproduce every shipped output exactly (the
reference prints correctly rounded doubles); in fp64 mode it must agree to
1e-11 except for the documented long-branch derivative cancellation.
"""
import pytest

from oracle import arbplf_oracle as O
from tests import helpers as H

CASES = H.manifest()
# keep the CPU suite short: the two 895-site GeLL cases take ~3 s each in mp mode
MP_CASES = [c for c in CASES]
FP64_SKIP = {"jc_long_deriv", "jc29_same_deriv", "jc29_diff_deriv", "jc30_same_deriv", "jc30_diff_deriv",
             "jc600_same_deriv"}


@pytest.mark.parametrize("case", MP_CASES, ids=[c["name"] for c in MP_CASES])
def test_oracle_mp_reproduces_golden_exactly(case):
    got = O.run(case["program"], H.golden_in(case["name"]), mode="mp")
    want = H.golden_out(case["name"])
    assert got["columns"] == want["columns"]
    assert len(got["data"]) == len(want["data"])
    for r1, r2 in zip(got["data"], want["data"]):
        assert r1[:-1] == r2[:-1]
        assert r1[-1] == r2[-1], (r1, r2)     # bit-for-bit


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_fp64_close_to_golden(case):
    if case["name"] in FP64_SKIP:
        pytest.skip("fp64 cancellation hazard documented in SURVEY.md section 7 (hard part 1)")
    got = O.run(case["program"], H.golden_in(case["name"]), mode="fp64")
    H.assert_tables_close(got, H.golden_out(case["name"]), rtol=1e-11, atol=1e-14, what=case["name"])


def test_yang_1994_gamma_rates():
    # test_scripts/test_gamma_discretization.py:101-107
    from mpmath import mp
    mp.prec = O.MP_PREC_BITS
    r = [float(x) for x in O.gamma_rates_mean(4, 0.5)]
    want = [0.0333877533835995, 0.251915917593438, 0.820268481973649, 2.89442784704931]
    for a, b in zip(r, want):
        assert abs(a - b) <= 1e-14 * b


def test_literal_derivative_matches_outside_identity():
    """arbplfderiv.c:112-207 restated literally == outside-pass identity."""
    prob = H.random_problem(3, ntips=7, n=4, S=4, ncat=2)
    a = O.run_deriv(prob, mode="mp", literal=True)
    b = O.run_deriv(prob, mode="mp", literal=False)
    H.assert_tables_close(a, b, rtol=1e-60, atol=1e-80)
