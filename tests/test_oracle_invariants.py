"""
The reference's invariant / closed-form tests (test_scripts/), run against the oracle on the CPU.  The goldens pin the
oracle on 4-state examples; these pin it on what the goldens do not cover -- other state counts, soft data, mixtures --
through properties that hold for any correct implementation.  The GPU path is compared with the oracle elsewhere
(tests/test_engine_gpu.py, tests/test_json_api_gpu.py), so the properties carry over.

  test_rate_mixture_vs_block.py        a rate mixture == the block-diagonal model over (category, state)
  test_em_monotonicity.py:95-131       EM updates of the edge rates never decrease the log-likelihood
  test_ll_deriv_accuracy.py:109,
  test_ll_hessian_accuracy.py:137      derivative and Hessian against finite differences
  test_marginal_no_change.py           pure-birth path 0--1--2 with both ends unborn: ll = -2 exactly, node 1 certain
  test_root_prior.py:98-140            the implicit equilibrium root prior == the same vector given explicitly
  test_path.py:133-195                 relabelling nodes / reordering edges permutes the outputs and nothing else
"""
import copy
import json

import numpy as np
import pytest

from oracle import arbplf_oracle as O
from tests import helpers as H


def _values(out):
    return np.array([r[-1] for r in out["data"]], dtype=np.float64)


def _block_model(md, rates, prior):
    """The (C n)-state model equivalent to `md` with the rate mixture (rates, prior): block-diagonal rate matrix, every
    observation repeated in each block, the category prior folded into the root prior."""
    Q = np.array(md["rate_matrix"], dtype=np.float64)
    n, C = Q.shape[0], len(rates)
    B = np.zeros((C * n, C * n))
    for c, r in enumerate(rates):
        B[c * n:(c + 1) * n, c * n:(c + 1) * n] = r * Q
    out = copy.deepcopy(md)
    out.pop("rate_mixture", None)
    out["rate_matrix"] = B.tolist()
    out["character_definitions"] = [list(d) * C for d in md["character_definitions"]]
    root = np.array(md["root_prior"], dtype=np.float64)
    out["root_prior"] = np.concatenate([p * root for p in prior]).tolist()
    return out


def _mixture_problem(seed, n=3, ntips=5, S=4):
    prob = H.random_problem(seed, ntips=ntips, n=n, S=S, ncat=1, root="custom", divisor=1.7, missing=0.2)
    md = prob["model_and_data"]
    for k in ("rate_mixture", "gamma_rate_mixture", "normalized_median_gamma_rate_mixture"):
        md.pop(k, None)
    return md


@pytest.mark.parametrize("seed", [51, 52])
def test_rate_mixture_equals_block_diagonal_model(seed):
    md = _mixture_problem(seed)
    rates, prior = [0.4, 1.0, 2.5], [0.2, 0.5, 0.3]
    mix = dict(md, rate_mixture={"rates": rates, "prior": prior})
    blk = _block_model(md, rates, prior)
    extra = {"site_reduction": {"aggregation": "sum"}}
    for run in (O.run_ll, O.run_deriv, O.run_em_update, O.run_hess):
        a = run(dict({"model_and_data": mix}, **extra), mode="fp64")
        b = run(dict({"model_and_data": blk}, **extra), mode="fp64")
        assert a["columns"] == b["columns"]
        assert [r[:-1] for r in a["data"]] == [r[:-1] for r in b["data"]]
        va, vb = _values(a), _values(b)
        assert np.allclose(va, vb, rtol=1e-10, atol=1e-12 * np.abs(vb).max()), run.__name__
    # per site as well (no aggregation), log-likelihood and derivatives
    for run in (O.run_ll, O.run_deriv):
        va = _values(run({"model_and_data": mix}, mode="fp64"))
        vb = _values(run({"model_and_data": blk}, mode="fp64"))
        assert np.allclose(va, vb, rtol=1e-10, atol=1e-12 * np.abs(vb).max())


def test_em_updates_never_decrease_the_log_likelihood():
    prob = H.random_problem(61, ntips=7, n=4, S=12, ncat=3, missing=0.1)
    md = prob["model_and_data"]
    doc = {"model_and_data": md, "site_reduction": {"aggregation": "sum"}}
    last = _values(O.run_ll(doc, mode="fp64"))[0]
    for _ in range(6):
        new = O.run_em_update(doc, mode="fp64")
        assert [r[0] for r in new["data"]] == list(range(len(md["edge_rate_coefficients"])))
        md["edge_rate_coefficients"] = [float(r[1]) for r in new["data"]]
        assert all(x >= 0.0 for x in md["edge_rate_coefficients"])
        ll = _values(O.run_ll(doc, mode="fp64"))[0]
        assert ll >= last - 1e-10 * abs(last)
        last = ll


def test_derivative_and_hessian_against_finite_differences():
    prob = H.random_problem(71, ntips=5, n=4, S=6, ncat=2, missing=0.1)
    md = prob["model_and_data"]
    doc = {"model_and_data": md, "site_reduction": {"aggregation": "sum"}}
    t0 = list(md["edge_rate_coefficients"])
    E = len(t0)
    grad = _values(O.run_deriv(doc, mode="fp64"))
    hess = _values(O.run_hess(doc, mode="fp64")).reshape(E, E)
    assert np.allclose(hess, hess.T, rtol=1e-10, atol=1e-12 * np.abs(hess).max())

    def at(t, run):
        md["edge_rate_coefficients"] = list(t)
        try:
            return _values(run(doc, mode="fp64"))
        finally:
            md["edge_rate_coefficients"] = list(t0)

    for e in range(E):
        h = 1e-5 * max(t0[e], 0.05)
        up, dn = list(t0), list(t0)
        up[e] += h
        dn[e] -= h
        fd = (at(up, O.run_ll)[0] - at(dn, O.run_ll)[0]) / (2 * h)
        assert abs(fd - grad[e]) <= 1e-6 * max(1.0, abs(grad[e]))
        fd_row = (at(up, O.run_deriv) - at(dn, O.run_deriv)) / (2 * h)
        assert np.allclose(fd_row, hess[e], rtol=1e-5, atol=1e-6 * max(1.0, np.abs(hess).max()))


def test_pure_birth_path_with_both_ends_unborn():
    md = {"edges": [[0, 1], [1, 2]], "edge_rate_coefficients": [1, 1], "rate_matrix": [[0, 1], [0, 0]],
          "probability_array": [[[1, 0], [1, 1], [1, 0]]]}
    ll = O.run_ll({"model_and_data": md}, mode="mp")
    assert ll["data"] == [[0, -2.0]]
    marg = O.run_marginal({"model_and_data": md}, mode="mp")
    got = {(r[-3], r[-2]): r[-1] for r in marg["data"]}
    for node in range(3):
        assert got[(node, 0)] == 1.0 and got[(node, 1)] == 0.0
    # heterogeneous rates and edges listed child first: -(t0 + t1)
    md2 = dict(md, edges=[[1, 2], [0, 1]], edge_rate_coefficients=[0.5, 2.25])
    assert O.run_ll({"model_and_data": md2}, mode="mp")["data"] == [[0, -2.75]]
    # no transition is ever expected, the whole of every edge is dwelt in state 0
    tr = O.run_trans({"model_and_data": md2, "trans_reduction": {"aggregation": "sum"}}, mode="mp")
    assert all(r[-1] == 0.0 for r in tr["data"])
    dw = O.run_dwell({"model_and_data": md2}, mode="mp")
    for r in dw["data"]:
        assert r[-1] == (1.0 if r[-2] == 0 else 0.0)


def test_implicit_equilibrium_root_prior_is_the_explicit_vector():
    prob = H.random_problem(81, ntips=6, n=4, S=5, ncat=2, root="equilibrium_distribution", missing=0.1)
    md = prob["model_and_data"]
    m = O.parse_model(md)
    be = O.get_backend("mp")
    cs = O.cross_site(m, be)
    pi = [float(x) for x in O.root_prior_vector(m, cs, be)]
    assert abs(sum(pi) - 1.0) < 1e-15
    explicit = dict(md, root_prior=pi)
    for run in (O.run_ll, O.run_deriv, O.run_marginal):
        a = _values(run({"model_and_data": md}, mode="fp64"))
        b = _values(run({"model_and_data": explicit}, mode="fp64"))
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


def test_relabelled_nodes_and_reordered_edges():
    prob = H.random_problem(91, ntips=6, n=4, S=5, ncat=2, missing=0.1, internal_data=True)
    md = prob["model_and_data"]
    N = len(md["character_data"][0])
    E = len(md["edges"])
    rng = np.random.default_rng(5)
    node_map = rng.permutation(N)                      # old label -> new label
    edge_order = rng.permutation(E)                    # new position -> old position
    new = copy.deepcopy(md)
    new["edges"] = [[int(node_map[md["edges"][k][0]]), int(node_map[md["edges"][k][1]])] for k in edge_order]
    new["edge_rate_coefficients"] = [md["edge_rate_coefficients"][k] for k in edge_order]
    inv = np.argsort(node_map)                         # new label -> old label
    new["character_data"] = [[row[int(inv[a])] for a in range(N)] for row in md["character_data"]]
    a = _values(O.run_ll({"model_and_data": md}, mode="fp64"))
    b = _values(O.run_ll({"model_and_data": new}, mode="fp64"))
    assert np.allclose(a, b, rtol=1e-12)
    S = len(md["character_data"])
    da = _values(O.run_deriv({"model_and_data": md}, mode="fp64")).reshape(S, E)
    db = _values(O.run_deriv({"model_and_data": new}, mode="fp64")).reshape(S, E)
    assert np.allclose(db, da[:, edge_order], rtol=1e-10, atol=1e-13)
    ma = _values(O.run_marginal({"model_and_data": md}, mode="fp64")).reshape(S, N, 4)
    mb = _values(O.run_marginal({"model_and_data": new}, mode="fp64")).reshape(S, N, 4)
    assert np.allclose(mb[:, node_map, :], ma, rtol=1e-10, atol=1e-13)


def _codon_problem(taxa=5, S=3, seed=3):
    """61-state GY94-style model (bench.codon_model) on a small tree with observed codons at the leaves."""
    import bench
    Q, pi = bench.codon_model()
    n = Q.shape[0]
    edges, N = bench.yule_tree(taxa, seed=seed)
    rng = np.random.default_rng(seed + 1)
    defs = np.vstack([np.eye(n), np.ones((1, n))])
    parents = {a for a, b in edges}
    data = [[int(rng.integers(0, n)) if a not in parents else n for a in range(N)] for _ in range(S)]
    return {"edges": edges, "edge_rate_coefficients": [float(x) for x in rng.exponential(0.1, len(edges))],
            "rate_matrix": Q.tolist(), "root_prior": "equilibrium_distribution", "rate_divisor": "equilibrium_exit_rate",
            "character_definitions": defs.tolist(), "character_data": data}, pi


def test_codon_model_equilibrium_and_finite_differences():
    """The cfg4 model family has no golden in the reference: the oracle's n-generic code is held to the model's
    known equilibrium (F3x4 frequencies: the rate matrix is reversible by construction) and to finite differences."""
    md, pi = _codon_problem()
    m = O.parse_model(md)
    be = O.get_backend("fp64")
    cs = O.cross_site(m, be)
    got_pi = np.array([float(x) for x in O.root_prior_vector(m, cs, be)])
    assert np.allclose(got_pi, pi, rtol=1e-11)
    doc = {"model_and_data": md, "site_reduction": {"aggregation": "sum"}}
    grad = _values(O.run_deriv(doc, mode="fp64"))
    t0 = list(md["edge_rate_coefficients"])
    for e in (0, len(t0) // 2, len(t0) - 1):
        h = 1e-5 * max(t0[e], 0.05)
        vals = []
        for sgn in (+1, -1):
            t = list(t0)
            t[e] += sgn * h
            md["edge_rate_coefficients"] = t
            vals.append(_values(O.run_ll(doc, mode="fp64"))[0])
        md["edge_rate_coefficients"] = list(t0)
        assert abs((vals[0] - vals[1]) / (2 * h) - grad[e]) <= 1e-6 * max(1.0, abs(grad[e]))
    # marginals are distributions, leaves with an observed codon are certain
    marg = O.run_marginal({"model_and_data": md}, mode="fp64")
    S, N = len(md["character_data"]), len(md["character_data"][0])
    mv = _values(marg).reshape(S, N, 61)
    assert np.allclose(mv.sum(axis=2), 1.0, atol=1e-12)
    for s in range(S):
        for a in range(N):
            code = md["character_data"][s][a]
            if code < 61:
                assert abs(mv[s, a, code] - 1.0) < 1e-12


def test_twenty_state_mixture_equals_forty_state_block_model():
    rng = np.random.default_rng(17)
    n = 20
    pi = rng.dirichlet(np.ones(n) * 3)
    R = rng.random((n, n)) + 0.1
    R = (R + R.T) / 2
    Q = R * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    md = _mixture_problem(53, n=4, ntips=5, S=3)
    N = len(md["character_data"][0])
    parents = {a for a, b in md["edges"]}
    md["rate_matrix"] = Q.tolist()
    md["root_prior"] = pi.tolist()
    md["character_definitions"] = np.vstack([np.eye(n), np.ones((1, n))]).tolist()
    md["character_data"] = [[int(rng.integers(0, n)) if a not in parents else n for a in range(N)] for _ in range(3)]
    rates, prior = [0.5, 2.0], [0.6, 0.4]
    mix = dict(md, rate_mixture={"rates": rates, "prior": prior})
    blk = _block_model(md, rates, prior)
    assert len(blk["rate_matrix"]) == 40
    for run in (O.run_ll, O.run_deriv):
        va = _values(run({"model_and_data": mix}, mode="fp64"))
        vb = _values(run({"model_and_data": blk}, mode="fp64"))
        assert np.allclose(va, vb, rtol=1e-10, atol=1e-12 * np.abs(vb).max())
