"""The C restatement (CPU baseline of bench.py) against the 320-bit Python oracle."""
import numpy as np
import pytest

from oracle import arbplf_oracle as O
from oracle import c_port
from tests import helpers as H
from tests.test_engine_gpu import PROBLEMS


@pytest.mark.parametrize("name,prob", PROBLEMS, ids=[p[0] for p in PROBLEMS])
def test_c_port_matches_oracle(name, prob):
    m = O.parse_model(prob["model_and_data"])
    ref = H.seam_reference(name, prob)
    defs, codes = H.dedupe_rows(m.dense_pmat())
    if codes.dtype != np.uint8:
        pytest.skip("C port takes 1-byte codes")
    params, cs = H.model_params(m)
    t = m.tree
    rng = np.random.default_rng(1)
    w = rng.random(m.site_count) + 0.5
    for threads in (1, 3):
        site_ll, sum_ll, sum_d = c_port.ll_deriv(t.indptr, t.indices, t.preorder, ref["P"], ref["Dm"], params["cat_prior"],
                                                 m.root_mode, params["root_vec"], codes, defs, w=w, nthreads=threads)
        np.testing.assert_allclose(site_ll, ref["ll"], rtol=1e-11, atol=2e-15)
        want = (w[:, None] * ref["D"]).sum(axis=0)
        floor = 2e-14 * (w[:, None] * ref["Dabs"]).sum(axis=0)
        assert np.all(np.abs(sum_d - want) <= 1e-11 * np.abs(want) + floor + 1e-300)
        assert abs(sum_ll - (w * ref["ll"]).sum()) <= 1e-11 * abs((w * ref["ll"]).sum()) + 1e-13
