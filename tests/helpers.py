"""Shared test helpers: fixtures from tests/golden, synthetic problem
generators, and glue that fills the device seam from the oracle's model
structures (tests are allowed to import the oracle; the product is not)."""
import json
import os

import numpy as np

from oracle import arbplf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def golden_in(name):
    with open(os.path.join(GOLDEN, name + ".in.json")) as f:
        return json.load(f)


def golden_out(name):
    with open(os.path.join(GOLDEN, name + ".out.json")) as f:
        return json.load(f)


def close(a, b, rtol=1e-11, atol=0.0):
    a = float(a)
    b = float(b)
    if a == b:
        return True
    return abs(a - b) <= rtol * abs(b) + atol


def assert_tables_close(got, want, rtol=1e-11, atol=0.0, what=""):
    assert got["columns"] == want["columns"], (what, got["columns"], want["columns"])
    assert len(got["data"]) == len(want["data"]), what
    for r1, r2 in zip(got["data"], want["data"]):
        assert r1[:-1] == r2[:-1], (what, r1, r2)
        assert close(r1[-1], r2[-1], rtol, atol), (what, r1, r2)


def dedupe_rows(dense):
    """dense [S,N,n] -> (defs [K,n], codes [S,N] uint8 or int32)."""
    S, N, n = dense.shape
    flat = np.ascontiguousarray(dense.reshape(S * N, n))
    defs, inv = np.unique(flat, axis=0, return_inverse=True)
    codes = inv.reshape(S, N)
    if defs.shape[0] <= 256:
        codes = codes.astype(np.uint8)
    else:
        codes = codes.astype(np.int32)
    return np.ascontiguousarray(defs), np.ascontiguousarray(codes)


def model_params(m):
    """Model constants for plf_set_model, from the oracle's 320-bit cross-site workspace."""
    from mpmath import mpf
    be = O.get_backend("mp")
    cs = O.cross_site(m, be, compute_P=False)
    n = m.n
    q_hi = np.zeros((n, n))
    q_lo = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            q_hi[i, j] = float(cs.Q[i, j])
            q_lo[i, j] = float(cs.Q[i, j] - mpf(q_hi[i, j]))
    edge_rates = np.array([float(x) for x in cs.edge_rates])
    rates = np.array([float(x) for x in cs.rates])
    prior = np.array([float(x) for x in cs.prior])
    root_vec = None
    if m.root_mode == O.ROOT_EQUILIBRIUM:
        root_vec = np.array([float(x) for x in cs.equilibrium])
    elif m.root_mode == O.ROOT_CUSTOM:
        root_vec = np.array(m.root_custom, dtype=np.float64)
    return dict(q_hi=q_hi, q_lo=q_lo, edge_rates=edge_rates, cat_rates=rates, cat_prior=prior,
                root_mode=m.root_mode, root_vec=root_vec), cs


def fill_engine(eng, m):
    t = m.tree
    eng.set_tree(t.indptr, t.indices, t.preorder)
    params, cs = model_params(m)
    eng.set_model(**params)
    defs, codes = dedupe_rows(m.dense_pmat())
    eng.set_data(defs, codes)
    return cs


def random_tree(rng, ntips, max_degree=2, root_degree=None):
    """Random rooted tree with arbitrary node labels; returns the edge list in random order."""
    # grow by splitting random leaves
    children = {0: []}
    leaves = [0]
    nxt = 1
    first = True
    while len(leaves) < ntips:
        a = leaves.pop(rng.integers(len(leaves)))
        deg = 2
        if first and root_degree:
            deg = root_degree
        elif max_degree > 2 and rng.random() < 0.3:
            deg = int(rng.integers(2, max_degree + 1))
        first = False
        deg = min(deg, ntips - len(leaves))
        deg = max(deg, 1)
        kids = list(range(nxt, nxt + deg))
        nxt += deg
        children[a] = kids
        for k in kids:
            children[k] = []
            leaves.append(k)
    N = nxt
    perm = rng.permutation(N)
    edges = [[int(perm[a]), int(perm[b])] for a in children for b in children[a]]
    order = rng.permutation(len(edges))
    edges = [edges[i] for i in order]
    leaf_labels = sorted(int(perm[l]) for l in leaves)
    return edges, N, leaf_labels


def random_problem(seed, ntips=6, n=4, S=5, ncat=1, root="custom", divisor="equilibrium_exit_rate",
                   missing=0.1, internal_data=False, max_degree=2, root_degree=None, soft=False,
                   mixture="custom", edge_scale=0.2):
    rng = np.random.default_rng(seed)
    edges, N, leaves = random_tree(rng, ntips, max_degree, root_degree)
    E = len(edges)
    Q = rng.random((n, n)) + 0.05
    np.fill_diagonal(Q, 0.0)
    md = {
        "edges": edges,
        "edge_rate_coefficients": [float(x) for x in rng.exponential(edge_scale, E)],
        "rate_matrix": [[float(x) for x in row] for row in Q],
    }
    if divisor is not None:
        md["rate_divisor"] = divisor
    if root == "custom":
        p = rng.random(n) + 0.1
        md["root_prior"] = [float(x) for x in p / p.sum()]
    elif root in ("equilibrium_distribution", "uniform_distribution"):
        md["root_prior"] = root
    if ncat > 1:
        if mixture == "custom":
            pr = rng.random(ncat) + 0.2
            md["rate_mixture"] = {"rates": [float(x) for x in rng.random(ncat) * 2], "prior": [float(x) for x in pr / pr.sum()]}
        elif mixture == "gamma":
            md["gamma_rate_mixture"] = {"gamma_shape": 0.7, "gamma_categories": ncat}
        elif mixture == "median_inv":
            md["normalized_median_gamma_rate_mixture"] = {"gamma_shape": 0.5, "gamma_categories": ncat - 1, "invariable_prior": 0.2}
    leafset = set(leaves)
    if soft:
        pa = rng.random((S, N, n)) + 0.01
        for s in range(S):
            for a in range(N):
                if a not in leafset and not internal_data:
                    pa[s, a, :] = 1.0
        md["probability_array"] = [[[float(x) for x in pa[s, a]] for a in range(N)] for s in range(S)]
    else:
        defs = [[1.0 if i == k else 0.0 for i in range(n)] for k in range(n)] + [[1.0] * n]
        data = []
        for s in range(S):
            row = []
            for a in range(N):
                if a in leafset or (internal_data and rng.random() < 0.3):
                    row.append(int(n if rng.random() < missing else rng.integers(n)))
                else:
                    row.append(n)
            data.append(row)
        md["character_definitions"] = defs
        md["character_data"] = data
    return {"model_and_data": md}


SEAM_CACHE = os.path.join(GOLDEN, "seam")


def _fl(a):
    return np.array(a.tolist(), dtype=np.float64) if a.dtype == object else np.asarray(a, dtype=np.float64)


def frechet_directions(m, cs):
    """Test directions: dwell weights on the diagonal, trans weights .* Q off it (hi/lo split)."""
    from mpmath import mpf
    n = m.n
    Ld = np.diag(np.linspace(0.25, 1.0, n))
    Lt = np.empty((n, n), dtype=object)
    for i in range(n):
        for j in range(n):
            Lt[i, j] = cs.Q[i, j] * (1 + ((i + 2 * j) % 3)) if i != j else mpf(0)
    Lt_hi = np.array([[float(x) for x in r] for r in Lt])
    Lt_lo = np.array([[float(Lt[i, j] - mpf(Lt_hi[i, j])) for j in range(n)] for i in range(n)])
    Lg = np.zeros((n, n))
    Lg[0, 0] = 1.0
    Lg[n - 1, 0] = 0.5
    return Ld, Lt, Lt_hi, Lt_lo, Lg


def seam_reference(name, prob):
    """
    Oracle (320-bit) values of everything the device seam returns for this
    problem, cached as tests/golden/seam/<name>.npz (generated by this function
    on first use; the cache is committed so that the GPU box does not spend
    minutes in mpmath).
    """
    path = os.path.join(SEAM_CACHE, name + ".npz")
    if os.path.exists(path):
        z = np.load(path)
        return {k: z[k] for k in z.files}
    m = O.parse_model(prob["model_and_data"])
    be = O.get_backend("mp")
    S = m.site_count
    sites = list(range(S))
    if name.startswith("deep"):
        sites = [0, 7, 39]
    ll, D, cs = O.per_site_ll_and_deriv(m, be, sites)
    out = {"sites": np.array(sites), "ll": _fl(ll), "D": _fl(D), "C": np.array(cs.C)}
    # magnitude of the cancelling terms of each derivative (fp64 is plenty for a scale)
    _, Dabs, _ = O.per_site_ll_and_deriv(m, be if name.startswith("deep") else O.get_backend("fp64"),
                                         sites, absQ=True)
    out["Dabs"] = _fl(Dabs)
    if not name.startswith("deep"):
        Ld, Lt, Lt_hi, Lt_lo, Lg = frechet_directions(m, cs)
        out["marg"] = _fl(O.per_site_marginal(m, be, sites))
        Xd, _ = O.per_site_edge_expect(m, be, sites, Ld, trans=False)
        Xt, _ = O.per_site_edge_expect(m, be, sites, Lt, trans=True)
        out["Xd"] = _fl(Xd)
        out["Xt"] = _fl(Xt)
        out["P"] = _fl(cs.P)
        Dm = np.empty(cs.P.shape, dtype=object)
        Dscale = np.zeros(cs.P.shape[:2])
        for c in range(cs.C):
            for e in range(cs.E):
                Dm[c, e] = (cs.Q @ cs.P[c, e]) * cs.rates[c]
                Dscale[c, e] = float(max((abs(cs.Q) @ cs.P[c, e]).flatten())) * float(cs.rates[c])
        out["Dm"] = _fl(Dm)
        out["Dscale"] = Dscale
        out["F"] = _fl(O.frechet_matrices(cs, be, be.asarray(Lg), [True] * cs.E))
    os.makedirs(SEAM_CACHE, exist_ok=True)
    np.savez_compressed(path, **out)
    return out


JSON_CACHE = os.path.join(GOLDEN, "json_cases")


def reduction_cases():
    """
    (name, program, document) triples exercising the selection / aggregation
    semantics of every axis (test_ll.py:64-94, test_ll_deriv.py:103-389,
    test_marginal.py:89-126, test_dwell.py, test_trans.py of the reference).
    """
    cases = []
    base = random_problem(31, ntips=7, n=4, S=6, ncat=2, root_degree=3)
    E = len(base["model_and_data"]["edges"])
    N = E + 1

    def doc(**red):
        d = {"model_and_data": base["model_and_data"]}
        d.update(red)
        return d
    cases += [
        ("ll_plain", "ll", doc()),
        ("ll_sum", "ll", doc(site_reduction={"aggregation": "sum"})),
        ("ll_avg", "ll", doc(site_reduction={"aggregation": "avg"})),
        ("ll_sel", "ll", doc(site_reduction={"selection": [4, 1, 1, 0]})),
        ("ll_sel_sum", "ll", doc(site_reduction={"selection": [4, 1, 1, 0], "aggregation": "sum"})),
        ("ll_sel_avg", "ll", doc(site_reduction={"selection": [4, 1, 1, 0], "aggregation": "avg"})),
        ("ll_only", "ll", doc(site_reduction={"selection": [3], "aggregation": "only"})),
        ("ll_weighted", "ll", doc(site_reduction={"aggregation": [1, 2.5, 0, -1, 3, 0.125]})),
        ("ll_sel_weighted", "ll", doc(site_reduction={"selection": [2, 2, 5], "aggregation": [1.5, -0.5, 2]})),
        ("deriv_plain", "deriv", doc()),
        ("deriv_site_sum", "deriv", doc(site_reduction={"aggregation": "sum"})),
        ("deriv_edge_sum", "deriv", doc(edge_reduction={"aggregation": "sum"})),
        ("deriv_both", "deriv", doc(site_reduction={"aggregation": "avg"}, edge_reduction={"aggregation": "sum"})),
        ("deriv_sel", "deriv", doc(site_reduction={"selection": [5, 0]}, edge_reduction={"selection": [E - 1, 0, 3]})),
        ("deriv_sel_w", "deriv", doc(site_reduction={"selection": [5, 0, 0], "aggregation": [1, 2, 3]},
                                     edge_reduction={"selection": [E - 1, 0, 3], "aggregation": [0.5, -1, 2]})),
        ("deriv_edge_only", "deriv", doc(edge_reduction={"selection": [2], "aggregation": "only"})),
        ("marg_plain", "marginal", doc()),
        ("marg_site_sum", "marginal", doc(site_reduction={"aggregation": "sum"})),
        ("marg_node_sel", "marginal", doc(node_reduction={"selection": [N - 1, 0, 2]})),
        ("marg_state_w", "marginal", doc(state_reduction={"aggregation": [1, 0, 2, 0.5]})),
        ("marg_all", "marginal", doc(site_reduction={"aggregation": "avg"}, node_reduction={"selection": [1, 3], "aggregation": "sum"},
                                     state_reduction={"selection": [0, 3], "aggregation": "sum"})),
        ("dwell_plain", "dwell", doc()),
        ("dwell_state_sum", "dwell", doc(state_reduction={"selection": [1, 3], "aggregation": "sum"})),
        ("dwell_state_w", "dwell", doc(state_reduction={"aggregation": [0.1, 0.2, 0.3, 0.4]}, site_reduction={"aggregation": "sum"})),
        ("dwell_state_sel", "dwell", doc(state_reduction={"selection": [2, 0]}, edge_reduction={"selection": [1, 4], "aggregation": "avg"})),
        ("dwell_only", "dwell", doc(state_reduction={"selection": [0], "aggregation": "only"}, edge_reduction={"aggregation": "sum"},
                                    site_reduction={"aggregation": "sum"})),
        ("trans_all_sum", "trans", doc(trans_reduction={"aggregation": "sum"})),
        ("trans_all_avg", "trans", doc(trans_reduction={"aggregation": "avg"}, site_reduction={"aggregation": "sum"})),
        ("trans_none", "trans", doc(site_reduction={"selection": [1]}, edge_reduction={"selection": [0, 2]})),
        ("trans_sel", "trans", doc(trans_reduction={"selection": [[0, 1], [2, 3], [0, 1]]}, site_reduction={"aggregation": "sum"})),
        ("trans_sel_w", "trans", doc(trans_reduction={"selection": [[0, 1], [2, 3], [3, 0]], "aggregation": [1, 2, -0.5]},
                                     edge_reduction={"aggregation": "sum"})),
        ("trans_only", "trans", doc(trans_reduction={"selection": [[1, 2]], "aggregation": "only"})),
    ]
    # other models through the whole stack
    gam = random_problem(32, ntips=9, n=4, S=20, ncat=4, mixture="gamma", missing=0.3, root="equilibrium_distribution")
    cases.append(("gamma_ll_sum", "ll", dict(gam, site_reduction={"aggregation": "sum"})))
    cases.append(("gamma_deriv_sum", "deriv", dict(gam, site_reduction={"aggregation": "sum"})))
    cases.append(("gamma_marg_avg", "marginal", dict(gam, site_reduction={"aggregation": "avg"})))
    inv = random_problem(33, ntips=6, n=4, S=8, ncat=5, mixture="median_inv", root="uniform_distribution")
    cases.append(("inv_deriv", "deriv", inv))
    cases.append(("inv_trans_sum", "trans", dict(inv, trans_reduction={"aggregation": "sum"}, site_reduction={"aggregation": "sum"})))
    n3 = random_problem(34, ntips=5, n=3, S=4, ncat=2, soft=True, internal_data=True, root=None, divisor=2.5)
    for prog in ("ll", "deriv", "marginal", "dwell", "trans"):
        cases.append(("soft3_" + prog, prog, n3))
    path = {"model_and_data": {
        "edges": [[0, 1], [1, 2], [2, 3]], "edge_rate_coefficients": [0.5, 0.0, 1.5],
        "rate_matrix": [[0, 1], [0, 0]], "root_prior": [1, 0],
        "probability_array": [[[1, 0], [1, 1], [1, 1], [0, 1]], [[1, 0], [1, 1], [1, 1], [1, 0]]]}}
    for prog in ("ll", "deriv", "marginal", "dwell", "trans"):
        cases.append(("absorbing_path_" + prog, prog, path))
    return cases


def expected_json(name, program, doc):
    """Oracle (320-bit) output for a JSON case, cached under tests/golden/json_cases/."""
    path = os.path.join(JSON_CACHE, name + ".json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    out = O.run(program, doc, mode="mp")
    os.makedirs(JSON_CACHE, exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f)
    return out
