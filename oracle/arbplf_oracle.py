"""
CPU oracle for the arbplf hot path -- TEST INFRASTRUCTURE ONLY.

This module is a from-scratch restatement, in numpy / mpmath, of the algorithm
that the reference (argriffing/phyly, `arbplf`) implements for the programs
arbplf-ll / -deriv / -marginal / -dwell / -trans.  It is the *checker* used by
`tests/`, by `__graft_entry__.smoke()` and by the `cpu_baseline` leg of
`bench.py`.  Nothing in the product path (phyly_b200/) may import it.

Parity status: PINNED.  The reference cannot be compiled or imported here (it
needs Arb/FLINT/jansson headers and the Python-2 C API), so the oracle is
pinned against the reference's own golden input/output pairs
(/root/reference/examples/**/in*.json -> out*.json, copied as fixtures under
tests/golden/ by tests/golden/make_golden.py) and against the known-answer /
closed-form values in its test scripts.  In 'mp' mode (mpmath, 320 bits) the
oracle reproduces the goldens to the last printed digit, because the reference
prints correctly rounded doubles.

Reference anchors (file:line under /root/reference/src):
  schema / validation ........ parsemodel.c:786-913, parsereduction.c:161-392
  CSR tree + BFS order ....... csr_graph.c:28-46,62-87,102-177,217-229;
                               util.c:369-408; model.c:23-45
  rate mixtures .............. rate_mixture.c:166-229,287-339;
                               gamma_discretization.c:208-370
  equilibrium ................ equilibrium.c:20-88
  cross-site workspace ....... cross_site_ws.c:150-242
  pruning (inside) ........... evaluate_site_lhood.c:6-63, util.c:241-301,
                               arb_mat_extras.c:35-113, model.c:282-350
  outside pass ............... evaluate_site_forward.c:31-105, model.c:225-280
  marginal ................... evaluate_site_marginal.c:6-21,
                               arbplfmarginal.c:111-264
  derivative ................. arbplfderiv.c:112-371 (root-path recomputation;
                               restated here through the outside-pass identity
                               and cross-checked against the literal form)
  Frechet / dwell / trans .... util.c:500-548, evaluate_site_frechet.c:4-42,
                               arbplfdwell.c:116-312, arbplftrans.c:115-346
  reductions / output ........ reduction.c:24-118, ndaccum.c:197-437
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

try:  # mpmath is only needed for the 'mp' arithmetic mode
    import mpmath
    from mpmath import mp, mpf
except Exception:  # pragma: no cover
    mpmath = None

MP_PREC_BITS = 320


class OracleError(RuntimeError):
    """Raised where the reference returns a non-zero retcode."""


# --------------------------------------------------------------------------
# JSON helpers that distinguish integers from reals like jansson does
# (parsemodel.c:592,771 json_is_integer; parsereduction.c:47)
# --------------------------------------------------------------------------

def _is_int(x) -> bool:
    return isinstance(x, int) and not isinstance(x, bool)


def _is_num(x) -> bool:
    return (isinstance(x, (int, float)) and not isinstance(x, bool))


def _exists(x) -> bool:
    return x is not None


def _strict_keys(obj, required: Sequence[str], optional: Sequence[str], what: str):
    """json_unpack_ex(..., JSON_STRICT, "{s:o, s?o ...}") semantics."""
    if not isinstance(obj, dict):
        raise OracleError("%s: expected an object" % what)
    for k in required:
        if k not in obj:
            raise OracleError("%s: object item not found: %s" % (what, k))
    allowed = set(required) | set(optional)
    for k in obj:
        if k not in allowed:
            raise OracleError("%s: unexpected key %r" % (what, k))


def _nonneg_array(root, desired_len: int, what: str) -> List[float]:
    # parsemodel.c:26-75
    if not isinstance(root, list):
        raise OracleError("%s: not an array" % what)
    if len(root) != desired_len:
        raise OracleError("%s: unexpected array length (actual: %d desired: %d)"
                          % (what, len(root), desired_len))
    out = []
    for x in root:
        if not _is_num(x):
            raise OracleError("%s: not a number" % what)
        if x < 0:
            raise OracleError("%s: array entries must be nonnegative" % what)
        out.append(float(x))
    return out


# --------------------------------------------------------------------------
# Tree (integer work; must be bit exact)
# --------------------------------------------------------------------------

@dataclass
class Tree:
    node_count: int
    edge_count: int
    root: int
    indptr: List[int]          # CSR, length N+1
    indices: List[int]         # CSR, length E (child node per csr edge)
    order: List[int]           # user edge -> csr idx   (csr_graph.c:28-46)
    preorder: List[int]        # BFS level order        (csr_graph.c:102-177)
    idx_to_a: List[int]        # csr idx -> parent node (util.c:369-389)
    b_to_idx: List[int]        # node -> csr idx of its parent edge, -1 at root
    pre_to_idx: List[int]      # csr edges in BFS node order (util.c:391-408)

    @property
    def idx_to_user_edge(self) -> List[int]:
        inv = [0] * self.edge_count
        for e, idx in enumerate(self.order):
            inv[idx] = e
        return inv


def build_tree(edges) -> Tree:
    """parsemodel.c:210-368 + csr_graph.c + model.c:23-45."""
    if not isinstance(edges, list):
        raise OracleError("_validate_edges: not an array")
    E = len(edges)
    N = E + 1
    in_deg = [0] * N
    out_deg = [0] * N
    pairs = []
    for e in edges:
        if not (isinstance(e, list) and len(e) == 2 and _is_int(e[0]) and _is_int(e[1])):
            raise OracleError("_validate_edge: expected [i, i]")
        a, b = e
        for idx in (a, b):
            if idx < 0 or idx >= N:
                raise OracleError("_validate_edges: node index out of range")
        if a == b:
            raise OracleError("_validate_edges: edges cannot be loops")
        out_deg[a] += 1
        in_deg[b] += 1
        pairs.append((a, b))
    roots = [i for i in range(N) if in_deg[i] == 0]
    if len(roots) != 1:
        raise OracleError("_validate_edges: exactly one node should have in-degree 0")
    root = roots[0]
    for i in range(N):
        if in_deg[i] > 1:
            raise OracleError("_validate_edges: the in-degree of each node must be 0 or 1")
    for i in range(N):
        if in_deg[i] + out_deg[i] < 1:
            raise OracleError("_validate_edges: node %d is not an endpoint of any edge" % i)
    indptr = [0] * (N + 1)
    for i in range(N):
        indptr[i + 1] = indptr[i] + out_deg[i]
    indices = [-1] * E
    fill = [0] * N
    for a, b in pairs:
        indices[indptr[a] + fill[a]] = b
        fill[a] += 1
    # user edge -> csr idx
    order = [0] * E
    fill = [0] * N
    for i, (a, b) in enumerate(pairs):
        idx = indptr[a] + fill[a]
        assert indices[idx] == b
        order[i] = idx
        fill[a] += 1
    # BFS by levels
    visited = [0] * N
    preorder = []
    level = [root]
    visited[root] = 1
    while level:
        nxt = []
        for a in level:
            preorder.append(a)
            for j in range(indptr[a], indptr[a + 1]):
                b = indices[j]
                if visited[b]:
                    raise OracleError("validate_edges: topo sort failed")
                visited[b] = 1
                nxt.append(b)
        level = nxt
    if len(preorder) != N:
        raise OracleError("validate_edges: the topo sort contains %d of the %d nodes"
                          % (len(preorder), N))
    idx_to_a = [0] * E
    b_to_idx = [-1] * N
    for a in range(N):
        for idx in range(indptr[a], indptr[a + 1]):
            idx_to_a[idx] = a
            b_to_idx[indices[idx]] = idx
    pre_to_idx = []
    for a in preorder:
        pre_to_idx.extend(range(indptr[a], indptr[a + 1]))
    return Tree(N, E, root, indptr, indices, order, preorder, idx_to_a, b_to_idx, pre_to_idx)


# --------------------------------------------------------------------------
# Model
# --------------------------------------------------------------------------

ROOT_NONE, ROOT_UNIFORM, ROOT_EQUILIBRIUM, ROOT_CUSTOM = range(4)
MIX_NONE, MIX_UNIFORM, MIX_CUSTOM, MIX_GAMMA, MIX_GAMMA_MEDIAN = range(5)


@dataclass
class Model:
    tree: Tree
    edge_rate_coefficients: List[float]          # user edge order
    rate_matrix: List[List[float]]               # raw user matrix
    n: int
    site_count: int
    # data: either dense pmat[site][node][state] or codes + definitions
    pmat: Optional[np.ndarray] = None            # float64 [S,N,n]
    codes: Optional[np.ndarray] = None           # int32  [S,N]
    defs: Optional[np.ndarray] = None            # float64 [K,n]
    use_equilibrium_rate_divisor: bool = False
    rate_divisor: float = 1.0
    root_mode: int = ROOT_NONE
    root_custom: Optional[List[float]] = None
    mix_mode: int = MIX_NONE
    mix_rates: Optional[List[float]] = None
    mix_prior: Optional[List[float]] = None
    gamma_shape: float = 1.0
    gamma_categories: int = 1
    invariable_prior: float = 0.0

    def category_count(self) -> int:
        # rate_mixture.c:257-284
        if self.mix_mode == MIX_NONE:
            return 1
        if self.mix_mode in (MIX_UNIFORM, MIX_CUSTOM):
            return len(self.mix_rates)
        return self.gamma_categories + (1 if self.invariable_prior else 0)

    def uses_equilibrium(self) -> bool:
        return self.root_mode == ROOT_EQUILIBRIUM or self.use_equilibrium_rate_divisor

    def dense_pmat(self) -> np.ndarray:
        if self.pmat is not None:
            return self.pmat
        return self.defs[self.codes]


def parse_model(md) -> Model:
    """validate_model_and_data, parsemodel.c:786-913."""
    _strict_keys(md,
                 ["edges", "edge_rate_coefficients", "rate_matrix"],
                 ["probability_array", "character_definitions", "character_data",
                  "rate_divisor", "root_prior", "rate_mixture", "gamma_rate_mixture",
                  "normalized_median_gamma_rate_mixture"], "model_and_data")
    g = md.get
    mixtures = sum(1 for k in ("rate_mixture", "gamma_rate_mixture",
                               "normalized_median_gamma_rate_mixture") if _exists(g(k)))
    if mixtures > 1:
        raise OracleError("error: conflicting rate mixture options")
    if _exists(g("probability_array")) and _exists(g("character_data")):
        raise OracleError("probability_array and character_data both specified")
    if _exists(g("probability_array")) and _exists(g("character_definitions")):
        raise OracleError("probability_array and character_definitions both specified")

    tree = build_tree(md["edges"])
    erc = _nonneg_array(md["edge_rate_coefficients"], tree.edge_count,
                        "_validate_edge_rate_coefficients")
    # rate matrix, parsemodel.c:390-455
    rm = md["rate_matrix"]
    if not isinstance(rm, list):
        raise OracleError("_validate_rate_matrix: not an array")
    n = len(rm)
    mat = []
    for row in rm:
        if not isinstance(row, list):
            raise OracleError("_validate_rate_matrix: this row is not an array")
        if len(row) != n:
            raise OracleError("_validate_rate_matrix: row length mismatch")
        r = []
        for y in row:
            if not _is_num(y):
                raise OracleError("_validate_rate_matrix: not a number")
            if y < 0:
                raise OracleError("_validate_rate_matrix: entries must be nonnegative")
            r.append(float(y))
        mat.append(r)
    m = Model(tree=tree, edge_rate_coefficients=erc, rate_matrix=mat, n=n, site_count=0)
    N = tree.node_count
    if _exists(g("probability_array")):
        pa = md["probability_array"]
        if not isinstance(pa, list):
            raise OracleError("_validate_probability_array: expected an array")
        S = len(pa)
        arr = np.zeros((S, N, n), dtype=np.float64)
        for i, x in enumerate(pa):
            if not isinstance(x, list):
                raise OracleError("_validate_probability_array: expected an array")
            if len(x) != N:
                raise OracleError("_validate_probability_array: failed to match the number of nodes")
            for j, y in enumerate(x):
                arr[i, j, :] = _nonneg_array(y, n, "_validate_probability_array")
        m.pmat = arr
        m.site_count = S
    elif _exists(g("character_data")):
        cd = md["character_data"]
        defs = g("character_definitions")
        if not isinstance(cd, list):
            raise OracleError("expected 'character_data' to be an array")
        if not isinstance(defs, list):
            raise OracleError("expected 'character_definitions' to be an array")
        K = len(defs)
        dtab = np.zeros((K, n), dtype=np.float64)
        for j, y in enumerate(defs):
            dtab[j, :] = _nonneg_array(y, n, "character_definitions")
        S = len(cd)
        codes = np.zeros((S, N), dtype=np.int32)
        for i, x in enumerate(cd):
            if not isinstance(x, list):
                raise OracleError("character_data: expected an array")
            if len(x) != N:
                raise OracleError("character_data: failed to match the number of nodes")
            for j, y in enumerate(x):
                if not _is_int(y):
                    raise OracleError("character indices must be integers")
                if y < 0:
                    raise OracleError("character indices must be non-negative")
                if y >= K:
                    raise OracleError("character indices must be less than the character count")
                codes[i, j] = y
        m.codes = codes
        m.defs = dtab
        m.site_count = S
    else:
        raise OracleError("either 'probability_array' or 'character_data' must be specified")

    # rate divisor, parsemodel.c:82-126
    rd = g("rate_divisor")
    if _exists(rd):
        if isinstance(rd, str):
            if rd == "equilibrium_exit_rate":
                m.use_equilibrium_rate_divisor = True
            else:
                raise OracleError("_validate_rate_divisor: bad string")
        elif _is_num(rd):
            if rd <= 0:
                raise OracleError("_validate_rate_divisor: must be positive")
            m.rate_divisor = float(rd)
        else:
            raise OracleError("_validate_rate_divisor: bad type")
    # root prior, parsemodel.c:129-188
    rp = g("root_prior")
    if not _exists(rp):
        m.root_mode = ROOT_NONE
    elif isinstance(rp, str):
        if rp == "equilibrium_distribution":
            m.root_mode = ROOT_EQUILIBRIUM
        elif rp == "uniform_distribution":
            m.root_mode = ROOT_UNIFORM
        else:
            raise OracleError("_validate_root_prior: bad string")
    else:
        m.root_mode = ROOT_CUSTOM
        m.root_custom = _nonneg_array(rp, n, "_validate_root_prior")
    # mixtures, parsemodel.c:631-783
    if _exists(g("gamma_rate_mixture")) or _exists(g("normalized_median_gamma_rate_mixture")):
        if _exists(g("gamma_rate_mixture")):
            m.mix_mode = MIX_GAMMA
            gm = md["gamma_rate_mixture"]
        else:
            m.mix_mode = MIX_GAMMA_MEDIAN
            gm = md["normalized_median_gamma_rate_mixture"]
        _strict_keys(gm, ["gamma_shape", "gamma_categories"], ["invariable_prior"],
                     "gamma_rate_mixture")
        ip = gm.get("invariable_prior")
        if ip is None:
            m.invariable_prior = 0.0
        elif _is_num(ip):
            m.invariable_prior = float(ip)
        else:
            raise OracleError("invariable_prior: not a number")
        if not _is_num(gm["gamma_shape"]):
            raise OracleError("gamma_shape: not a number")
        m.gamma_shape = float(gm["gamma_shape"])
        if not _is_int(gm["gamma_categories"]):
            raise OracleError("gamma_categories: not an integer")
        m.gamma_categories = int(gm["gamma_categories"])
    elif _exists(g("rate_mixture")):
        rmx = md["rate_mixture"]
        _strict_keys(rmx, ["rates", "prior"], [], "rate_mixture")
        if not isinstance(rmx["rates"], list):
            raise OracleError("_validate_rate_mixture: 'rates' is not an array")
        k = len(rmx["rates"])
        m.mix_rates = _nonneg_array(rmx["rates"], k, "rate_mixture rates")
        pr = rmx["prior"]
        if isinstance(pr, str):
            if pr == "uniform_distribution":
                m.mix_mode = MIX_UNIFORM
            else:
                raise OracleError("_validate_rate_mixture: bad prior string")
        elif isinstance(pr, list):
            m.mix_prior = _nonneg_array(pr, k, "rate_mixture prior")
            m.mix_mode = MIX_CUSTOM
        else:
            # the reference leaves the mode undefined and aborts later
            raise OracleError("_validate_rate_mixture: bad prior")
    else:
        m.mix_mode = MIX_NONE
    return m


# --------------------------------------------------------------------------
# Reductions (parsereduction.c, reduction.c)
# --------------------------------------------------------------------------

AGG_NONE, AGG_AVG, AGG_SUM, AGG_WEIGHTED_SUM, AGG_ONLY = range(5)


@dataclass
class Reduction:
    selection: List[int]
    agg_mode: int = AGG_NONE
    weights: Optional[List[float]] = None
    # only for pair reductions
    first_idx: Optional[List[int]] = None
    second_idx: Optional[List[int]] = None


def _parse_aggregation(r: Reduction, name: str, agg, present: bool):
    # parsereduction.c:77-159
    if not present:
        r.agg_mode = AGG_NONE
    elif isinstance(agg, str):
        if agg == "sum":
            r.agg_mode = AGG_SUM
        elif agg == "avg":
            r.agg_mode = AGG_AVG
        elif agg == "only":
            if len(r.selection) != 1:
                raise OracleError("%s aggregation (only): selection length must be 1" % name)
            r.agg_mode = AGG_ONLY
        else:
            raise OracleError("%s aggregation: invalid string" % name)
    elif isinstance(agg, list):
        r.agg_mode = AGG_WEIGHTED_SUM
        if len(agg) != len(r.selection):
            raise OracleError("%s aggregation: weight count mismatch" % name)
        ws = []
        for x in agg:
            if not _is_num(x):
                raise OracleError("%s aggregation: weights should be numeric" % name)
            ws.append(float(x))
        r.weights = ws
    else:
        raise OracleError("%s aggregation: invalid type" % name)


def parse_column_reduction(root, k: int, name: str) -> Reduction:
    # parsereduction.c:161-195
    sel_present = agg_present = False
    sel = agg = None
    if root is not None:
        _strict_keys(root, [], ["selection", "aggregation"], name + "_reduction")
        sel_present = "selection" in root
        agg_present = "aggregation" in root
        sel = root.get("selection")
        agg = root.get("aggregation")
    if not sel_present:
        selection = list(range(k))
    else:
        if not isinstance(sel, list):
            raise OracleError("%s selection: should be an array" % name)
        selection = []
        for x in sel:
            if not _is_int(x):
                raise OracleError("%s selection: must be an integer" % name)
            if x < 0:
                raise OracleError("%s selection: must be non-negative" % name)
            if x >= k:
                raise OracleError("%s selection: out of range" % name)
            selection.append(x)
    r = Reduction(selection=selection)
    _parse_aggregation(r, name, agg, agg_present)
    return r


def parse_pair_reduction(root, k: int, name: str) -> Reduction:
    # parsereduction.c:290-392
    sel = agg = None
    agg_present = False
    if root is not None:
        _strict_keys(root, [], ["selection", "aggregation"], name + "_reduction")
        sel = root.get("selection")
        agg = root.get("aggregation")
        agg_present = "aggregation" in root
    if _exists(sel):
        if not isinstance(sel, list):
            raise OracleError("%s selection: should be an array" % name)
        first, second = [], []
        for p in sel:
            if not (isinstance(p, list) and len(p) == 2 and _is_int(p[0]) and _is_int(p[1])):
                raise OracleError("%s selection: expected [i, i]" % name)
            for idx in p:
                if idx < 0 or idx >= k:
                    raise OracleError("%s selection: index out of range" % name)
            first.append(p[0])
            second.append(p[1])
        r = Reduction(selection=list(range(len(first))), first_idx=first, second_idx=second)
        _parse_aggregation(r, name, agg, agg_present)
        return r
    mode = -1
    if _exists(agg):
        if agg == "sum":
            mode = AGG_SUM
        elif agg == "avg":
            mode = AGG_AVG
    else:
        mode = AGG_NONE
    if mode == -1:
        raise OracleError("%s reduction (no selection): only sum or avg allowed" % name)
    first, second = [], []
    for a in range(k):
        for b in range(k):
            if a != b:
                first.append(a)
                second.append(b)
    return Reduction(selection=list(range(len(first))), agg_mode=mode,
                     first_idx=first, second_idx=second)


# --------------------------------------------------------------------------
# Arithmetic back ends
# --------------------------------------------------------------------------

class _FP64:
    name = "fp64"

    def num(self, x):
        return float(x)

    def zeros(self, *shape):
        return np.zeros(shape, dtype=np.float64)

    def ones(self, *shape):
        return np.ones(shape, dtype=np.float64)

    def asarray(self, a):
        return np.asarray(a, dtype=np.float64)

    def log(self, x):
        return math.log(x)

    def expm(self, A):
        import scipy.linalg
        return scipy.linalg.expm(np.asarray(A, dtype=np.float64))

    def solve(self, A, b):
        return np.linalg.solve(np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64))

    def to_float(self, x):
        return float(x)


class _MP:
    name = "mp"

    def __init__(self, prec=MP_PREC_BITS):
        self.prec = prec

    def num(self, x):
        return mpf(x)

    def zeros(self, *shape):
        a = np.empty(shape, dtype=object)
        a.fill(mpf(0))
        return a

    def ones(self, *shape):
        a = np.empty(shape, dtype=object)
        a.fill(mpf(1))
        return a

    def asarray(self, a):
        a = np.asarray(a)
        out = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(a.shape):
            out[idx] = mpf(a[idx])
        return out

    def log(self, x):
        return mpmath.log(x)

    def expm(self, A):
        A = np.asarray(A, dtype=object)
        n = A.shape[0]
        M = mpmath.matrix(n, n)
        for i in range(n):
            for j in range(n):
                M[i, j] = A[i, j]
        R = mpmath.expm(M, method="taylor")
        out = np.empty((n, n), dtype=object)
        for i in range(n):
            for j in range(n):
                out[i, j] = R[i, j]
        return out

    def solve(self, A, b):
        A = np.asarray(A, dtype=object)
        n = A.shape[0]
        M = mpmath.matrix(n, n)
        v = mpmath.matrix(n, 1)
        for i in range(n):
            v[i] = b[i]
            for j in range(n):
                M[i, j] = A[i, j]
        x = mpmath.lu_solve(M, v)
        out = np.empty((n,), dtype=object)
        for i in range(n):
            out[i] = x[i]
        return out

    def to_float(self, x):
        # Values below the working-precision noise floor are exact zeros in the
        # reference (its precision loop runs until |mid|+rad < 2^-1075,
        # util.c:28-36); at 320 bits anything under 2^-272 is rounding noise.
        if abs(x) < mpf(2) ** (-(self.prec - 48)):
            return 0.0
        return float(x)


def get_backend(mode: str):
    if mode == "fp64":
        return _FP64()
    if mode == "mp":
        if mpmath is None:
            raise RuntimeError("mpmath not available")
        mp.prec = MP_PREC_BITS
        return _MP()
    raise ValueError(mode)


# --------------------------------------------------------------------------
# Gamma discretisation (gamma_discretization.c:208-370), always done in mp
# --------------------------------------------------------------------------

def _mp_gamma_P(s, x):
    if x == mpmath.inf:
        return mpf(1)
    if x == 0:
        return mpf(0)
    return mpmath.gammainc(s, 0, x, regularized=True)


def gamma_quantile(k: int, n: int, s) -> "mpf":
    """Quantile q with P(s, q) = k/n (regularised lower incomplete gamma)."""
    c = mpf(k) / n
    s = mpf(s)
    # work in log space: q = exp(u); robust for tiny shapes
    # initial guess from the small-x asymptotic P(s,x) ~ x^s / Gamma(s+1)
    lg = mpmath.log(c) + mpmath.loggamma(s + 1)
    u0 = lg / s
    f = lambda u: _mp_gamma_P(s, mpmath.exp(u)) - c
    # bracket
    lo = u0 - 1
    while f(lo) > 0:
        lo -= max(1, abs(lo))
    hi = max(u0, mpf(0)) + 1
    while f(hi) < 0:
        hi += max(1, abs(hi))
    # bisection + secant polish (findroot with a bracket: 'anderson')
    u = mpmath.findroot(f, (lo, hi), solver="anderson", tol=mpf(2) ** (-(mp.prec - 20)),
                        maxsteps=2000, verify=False)
    return mpmath.exp(u)


def gamma_rates_mean(n: int, s) -> List["mpf"]:
    # gamma_discretization.c:298-332
    s = mpf(s)
    q = [mpf(0)] + [gamma_quantile(k, n, s) for k in range(1, n)] + [mpmath.inf]
    ex = [_mp_gamma_P(s + 1, x) for x in q]
    return [(ex[k + 1] - ex[k]) * n for k in range(n)]


def gamma_rates_median(n: int, s) -> List["mpf"]:
    # gamma_discretization.c:334-370
    s = mpf(s)
    q = [gamma_quantile(2 * k + 1, 2 * n, s) for k in range(n)]
    tot = sum(q)
    return [x / tot * n for x in q]


def rate_mixture_summary(m: Model):
    """(prior[C], rates[C], expect) as mpf; rate_mixture.c:166-229,287-339."""
    old = mp.prec
    mp.prec = MP_PREC_BITS
    try:
        if m.mix_mode == MIX_NONE:
            return [mpf(1)], [mpf(1)], mpf(1)
        if m.mix_mode == MIX_UNIFORM:
            k = len(m.mix_rates)
            rates = [mpf(r) for r in m.mix_rates]
            return [mpf(1) / k] * k, rates, sum(rates) / k
        if m.mix_mode == MIX_CUSTOM:
            rates = [mpf(r) for r in m.mix_rates]
            prior = [mpf(p) for p in m.mix_prior]
            return prior, rates, sum(r * p for r, p in zip(rates, prior))
        K = m.gamma_categories
        p = mpf(m.invariable_prior)
        q = 1 - p
        if m.mix_mode == MIX_GAMMA:
            rates = gamma_rates_mean(K, m.gamma_shape)
        else:
            rates = gamma_rates_median(K, m.gamma_shape)
        rates = [r / q for r in rates]
        prior = [q / K] * K
        if m.invariable_prior:
            prior.append(p)
            rates.append(mpf(0))
        return prior, rates, mpf(1)
    finally:
        mp.prec = old


# --------------------------------------------------------------------------
# Cross-site workspace (cross_site_ws.c:199-242)
# --------------------------------------------------------------------------

@dataclass
class CrossSite:
    n: int
    C: int
    E: int
    prior: Any            # [C]
    rates: Any            # [C]
    expect: Any
    equilibrium: Any      # [n] or None
    rate_divisor: Any
    Q: Any                # [n,n] scaled, with diagonal
    edge_rates: Any       # [E] csr order
    P: Any                # [C,E,n,n]
    P_is_identity: Any = None


def cross_site(m: Model, be, edge_rates_user: Optional[Sequence[float]] = None,
               compute_P: bool = True) -> CrossSite:
    n = m.n
    t = m.tree
    prior_mp, rates_mp, expect_mp = rate_mixture_summary(m)
    if be.name == "fp64":
        prior = np.array([float(x) for x in prior_mp])
        rates = np.array([float(x) for x in rates_mp])
        expect = float(expect_mp)
    else:
        prior = np.array(prior_mp, dtype=object)
        rates = np.array(rates_mp, dtype=object)
        expect = expect_mp
    C = len(prior)
    raw = be.asarray(m.rate_matrix)
    for i in range(n):
        raw[i, i] = be.num(0)
    eq = None
    if m.uses_equilibrium():
        # equilibrium.c:20-88: [Q^T e; e^T 0] x = 1
        R = be.zeros(n + 1, n + 1)
        for i in range(n):
            ex = be.num(0)
            for j in range(n):
                if i != j:
                    R[i, j] = raw[j, i]
                    ex = ex + raw[i, j]
            R[i, i] = -ex
            R[n, i] = be.num(1)
            R[i, n] = be.num(1)
        b = be.ones(n + 1)
        x = be.solve(R, b)
        eq = x[:n]
    if m.use_equilibrium_rate_divisor:
        row_sums = raw.sum(axis=1)
        div = sum(row_sums[i] * eq[i] for i in range(n)) * expect
    else:
        div = be.num(m.rate_divisor)
    Q = raw / div
    for i in range(n):
        Q[i, i] = be.num(0)
        Q[i, i] = -sum(Q[i, j] for j in range(n) if j != i)
    er_user = m.edge_rate_coefficients if edge_rates_user is None else edge_rates_user
    er = be.zeros(t.edge_count)
    for i in range(t.edge_count):
        er[t.order[i]] = be.num(er_user[i])
    P = np.empty((C, t.edge_count, n, n), dtype=(np.float64 if be.name == "fp64" else object))
    for c in range(C if compute_P else 0):
        for e in range(t.edge_count):
            s = rates[c] * er[e]
            if s == 0:
                P[c, e] = be.asarray(np.eye(n))
            else:
                P[c, e] = be.expm(Q * s)
    return CrossSite(n=n, C=C, E=t.edge_count, prior=prior, rates=rates, expect=expect,
                     equilibrium=eq, rate_divisor=div, Q=Q, edge_rates=er, P=P)


def frechet_matrices(cs: CrossSite, be, L, edge_requested_csr: Sequence[bool]):
    """F[c][e] = top-right block of exp([[A,L],[0,A]]), A = rate_c t_e Q (util.c:500-548)."""
    n = cs.n
    F = np.empty((cs.C, cs.E, n, n), dtype=(np.float64 if be.name == "fp64" else object))
    for c in range(cs.C):
        for e in range(cs.E):
            if not edge_requested_csr[e]:
                F[c, e] = be.zeros(n, n)
                continue
            A = cs.Q * (cs.rates[c] * cs.edge_rates[e])
            M = be.zeros(2 * n, 2 * n)
            M[:n, :n] = A
            M[n:, n:] = A
            M[:n, n:] = L
            X = be.expm(M)
            F[c, e] = X[:n, n:]
    return F


# --------------------------------------------------------------------------
# Per-site kernels, vectorised over sites: arrays are [S, n]
# --------------------------------------------------------------------------

def _row_is_const(v) -> np.ndarray:
    """[S] bool: all entries of each row are equal (arb_mat_extras.c:35-51)."""
    return np.all(v == v[:, :1], axis=1)


def _matvec(P, v):
    """(P @ v_s) for each site s; v is [S,n] -> [S,n]."""
    return v @ P.T


def root_prior_vector(m: Model, cs: CrossSite, be):
    n = m.n
    if m.root_mode == ROOT_NONE:
        return be.ones(n)
    if m.root_mode == ROOT_UNIFORM:
        return be.ones(n) / be.num(n)
    if m.root_mode == ROOT_EQUILIBRIUM:
        return cs.equilibrium.copy()
    return be.asarray(m.root_custom)


def site_lhood(m: Model, cs: CrossSite, be, base, cat: int, keep_edges=False):
    """
    evaluate_site_lhood.c:6-63 for all sites at once.
    base: [S,N,n].  Returns (lhood[S], node_vecs{a:[S,n]}, node_const{a:[S] bool},
                             edge_vecs{idx:[S,n]} or None)
    The boolean 'const' flags carry the reference's exact-constant-column
    shortcut (util.c:276-283, arb_mat_extras.c:84-91) symbolically.
    """
    t = m.tree
    S = base.shape[0]
    node = {}
    nconst = {}
    edge = {} if keep_edges else None
    for a in reversed(t.preorder):
        v = base[:, a, :].copy()
        vc = _row_is_const(v)
        for idx in range(t.indptr[a], t.indptr[a + 1]):
            b = t.indices[idx]
            Pm = cs.P[cat, idx]
            em = _matvec(Pm, node[b])
            bc = nconst[b]
            if bc.any():
                em[bc] = node[b][bc]
            if keep_edges:
                edge[idx] = em
            v = v * em
            vc = vc & bc
        node[a] = v
        nconst[a] = vc
    rootv = node[t.root]
    rc = nconst[t.root]
    rp = root_prior_vector(m, cs, be)
    if m.root_mode == ROOT_NONE:
        lh = rootv.sum(axis=1)
    else:
        lh = rootv @ rp
        if m.root_mode in (ROOT_UNIFORM, ROOT_EQUILIBRIUM) and rc.any():
            # model.c:306-309, 319-323
            lh = lh.copy()
            lh[rc] = rootv[rc, 0]
    return lh, node, nconst, edge


def site_forward(m: Model, cs: CrossSite, be, base, cat: int, edge_vecs):
    """evaluate_site_forward.c:31-105.  Returns (fwd_node{a}, fwd_edge{idx})."""
    t = m.tree
    S = base.shape[0]
    n = m.n
    fn = {}
    fe = {}
    rp = root_prior_vector(m, cs, be)
    fn[t.root] = np.tile(rp, (S, 1))
    for a in t.preorder:
        tmp = fn[a] * base[:, a, :]
        start, stop = t.indptr[a], t.indptr[a + 1]
        for idx in range(start, stop):
            b = t.indices[idx]
            f = tmp.copy()
            for idx2 in range(start, stop):
                if idx2 != idx:
                    f = f * edge_vecs[idx2]
            fe[idx] = f
            fn[b] = f @ cs.P[cat, idx]      # P^T f
    return fn, fe


def site_derivatives_literal(m: Model, cs: CrossSite, be, base, cat: int,
                             edge_vecs, node_const, requested_csr):
    """
    Literal restatement of evaluate_site_derivatives (arbplfderiv.c:112-207):
    for each requested edge walk the path to the root.  O(E*depth); used to
    validate the outside-pass identity on small inputs.
    """
    t = m.tree
    S = base.shape[0]
    out = {}
    rp = root_prior_vector(m, cs, be)
    for d in range(t.edge_count):
        if not requested_csr[d]:
            continue
        dn = {}
        curr = d
        while curr != -1:
            a = t.idx_to_a[curr]
            v = base[:, a, :].copy()
            for idx in range(t.indptr[a], t.indptr[a + 1]):
                if idx == d:
                    em = edge_vecs[idx]
                    r = _matvec(cs.Q, em)
                    ec = _row_is_const(em) & node_const[t.indices[idx]]
                    if ec.any():
                        r[ec] = 0
                    v = v * r
                elif idx == curr:
                    b = t.indices[idx]
                    v = v * _matvec(cs.P[cat, idx], dn[b])
                else:
                    v = v * edge_vecs[idx]
            dn[a] = v
            curr = t.b_to_idx[a]
        rootv = dn[t.root]
        if m.root_mode == ROOT_NONE:
            out[d] = rootv.sum(axis=1)
        else:
            out[d] = rootv @ rp
    return out


# --------------------------------------------------------------------------
# nd accumulator (ndaccum.c) -- semantic restatement
# --------------------------------------------------------------------------

def agg_weights(r: Reduction, total_len: int, be):
    """get_column_agg_weights, reduction.c:24-118 -> (weights[total_len], divisor)."""
    w = [be.num(0)] * total_len
    w = list(w)
    if r.agg_mode == AGG_WEIGHTED_SUM:
        for i, idx in enumerate(r.selection):
            w[idx] = w[idx] + be.num(r.weights[i])
        return w, be.num(1)
    if r.agg_mode in (AGG_SUM, AGG_AVG):
        cnt = [0] * total_len
        for idx in r.selection:
            cnt[idx] += 1
        w = [be.num(c) for c in cnt]
        return w, (be.num(1) if r.agg_mode == AGG_SUM else be.num(len(r.selection)))
    if r.agg_mode == AGG_ONLY:
        if len(r.selection) != 1:
            raise OracleError("aggregation only: selection length must be 1")
        w[r.selection[0]] = be.num(1)
        return w, be.num(1)
    raise OracleError("internal error: unexpected aggregation mode")


@dataclass
class Axis:
    name: str
    n: int
    r: Reduction
    component_names: Optional[List[str]] = None
    component_indices: Optional[List[List[int]]] = None
    weights: Any = None
    divisor: Any = None

    @property
    def aggregated(self):
        return self.r.agg_mode != AGG_NONE

    def requested(self):
        req = [False] * self.n
        for idx in self.r.selection:
            req[idx] = True
        return req


class NdAccum:
    def __init__(self, axes: List[Axis], be):
        self.axes = axes
        self.be = be
        for ax in axes:
            if ax.aggregated:
                ax.weights, ax.divisor = agg_weights(ax.r, ax.n, be)
        self.shape = [1 if ax.aggregated else ax.n for ax in axes]
        self.data = {}

    def accumulate(self, coords, value):
        x = value
        key = []
        for ax, c in zip(self.axes, coords):
            if ax.aggregated:
                x = x * ax.weights[c]
                x = x / ax.divisor
                key.append(0)
            else:
                key.append(c)
        key = tuple(key)
        self.data[key] = self.data.get(key, self.be.num(0)) + x

    def to_json(self):
        cols = []
        for ax in self.axes:
            if not ax.aggregated:
                if ax.component_names:
                    cols.extend(ax.component_names)
                else:
                    cols.append(ax.name)
        cols.append("value")
        rows = []

        def rec(i, prefix, key):
            if i == len(self.axes):
                v = self.data.get(tuple(key), self.be.num(0))
                d = self.be.to_float(v)
                if d == 0.0:
                    d = 0.0  # scrub -0.0 (util.c:44-48)
                rows.append(prefix + [d])
                return
            ax = self.axes[i]
            if ax.aggregated:
                rec(i + 1, prefix, key + [0])
            else:
                for k, idx in enumerate(ax.r.selection):
                    if ax.component_indices:
                        p = prefix + [comp[k] for comp in ax.component_indices]
                    else:
                        p = prefix + [idx]
                    rec(i + 1, p, key + [idx])
        rec(0, [], [])
        return {"columns": cols, "data": rows}


# --------------------------------------------------------------------------
# Program drivers
# --------------------------------------------------------------------------

def _base_vectors(m: Model, be, sites: Sequence[int]):
    dense = m.dense_pmat()[list(sites)]
    if be.name == "fp64":
        return np.ascontiguousarray(dense, dtype=np.float64)
    return be.asarray(dense)


def _top(root, optional_axes: Sequence[str], what: str):
    _strict_keys(root, ["model_and_data"], list(optional_axes), what)
    return parse_model(root["model_and_data"])


def _requested_sites(site_r: Reduction, S: int) -> List[int]:
    req = [False] * S
    for i in site_r.selection:
        req[i] = True
    return [i for i in range(S) if req[i]]


def _site_likelihoods(m, cs, be, base, keep=False):
    """Per category inside pass.  Returns (site_L[S], per-cat results)."""
    S = base.shape[0]
    tot = be.zeros(S)
    per = []
    for c in range(cs.C):
        lh, node, nconst, edge = site_lhood(m, cs, be, base, c, keep_edges=keep)
        tot = tot + cs.prior[c] * lh
        per.append((lh, node, nconst, edge))
    return tot, per


def run_ll(root, mode="mp"):
    be = get_backend(mode)
    m = _top(root, ["site_reduction"], "arbplf-ll")
    r_site = parse_column_reduction(root.get("site_reduction"), m.site_count, "site")
    cs = cross_site(m, be)
    acc = NdAccum([Axis("site", m.site_count, r_site)], be)
    sites = _requested_sites(r_site, m.site_count)
    if sites:
        base = _base_vectors(m, be, sites)
        tot, _ = _site_likelihoods(m, cs, be, base)
        for k, s in enumerate(sites):
            if not (tot[k] > 0):
                raise OracleError("site %d has zero likelihood" % s)
            acc.accumulate([s], be.log(tot[k]))
    return acc.to_json()


def per_site_ll_and_deriv(m: Model, be, sites: Sequence[int], want_deriv=True,
                          requested_csr=None, literal=False, absQ=False):
    """
    Returns (ll[S'], deriv[S', E] in csr edge order).  Used by the tests to
    check the device seam directly.  With absQ=True the derivative sum is
    formed with |Q| instead of Q: the magnitude of the terms that cancel, i.e.
    the scale against which a floating-point evaluation can be accurate.
    """
    cs = cross_site(m, be)
    t = m.tree
    base = _base_vectors(m, be, sites)
    S = base.shape[0]
    if requested_csr is None:
        requested_csr = [True] * t.edge_count
    tot, per = _site_likelihoods(m, cs, be, base, keep=want_deriv)
    ll = np.array([be.log(x) for x in tot], dtype=(np.float64 if be.name == "fp64" else object))
    if not want_deriv:
        return ll, None, cs
    D = be.zeros(S, t.edge_count)
    for c in range(cs.C):
        lh, node, nconst, edge = per[c]
        if literal:
            dd = site_derivatives_literal(m, cs, be, base, c, edge, nconst, requested_csr)
            for idx, v in dd.items():
                D[:, idx] = D[:, idx] + cs.prior[c] * cs.rates[c] * v
        else:
            fn, fe = site_forward(m, cs, be, base, c, edge)
            for idx in range(t.edge_count):
                if not requested_csr[idx]:
                    continue
                b = t.indices[idx]
                qe = _matvec(abs(cs.Q) if absQ else cs.Q, edge[idx])
                ec = nconst[b]
                if ec.any() and not absQ:
                    qe[ec] = 0
                v = (fe[idx] * qe).sum(axis=1)
                D[:, idx] = D[:, idx] + cs.prior[c] * cs.rates[c] * v
    D = D / tot[:, None]
    return ll, D, cs


def run_deriv(root, mode="mp", literal=False):
    be = get_backend(mode)
    m = _top(root, ["site_reduction", "edge_reduction"], "arbplf-deriv")
    t = m.tree
    r_site = parse_column_reduction(root.get("site_reduction"), m.site_count, "site")
    r_edge = parse_column_reduction(root.get("edge_reduction"), t.edge_count, "edge")
    ax_edge = Axis("edge", t.edge_count, r_edge)
    acc = NdAccum([Axis("site", m.site_count, r_site), ax_edge], be)
    req_user = ax_edge.requested()
    req_csr = [False] * t.edge_count
    for e in range(t.edge_count):
        req_csr[t.order[e]] = req_user[e]
    sites = _requested_sites(r_site, m.site_count)
    if sites:
        _, D, _ = per_site_ll_and_deriv(m, be, sites, True, req_csr, literal=literal)
        for k, s in enumerate(sites):
            for idx in range(t.edge_count):
                if req_csr[idx]:
                    acc.accumulate([s, t.idx_to_user_edge[idx]], D[k, idx])
    return acc.to_json()


def run_marginal(root, mode="mp"):
    be = get_backend(mode)
    m = _top(root, ["site_reduction", "node_reduction", "state_reduction"], "arbplf-marginal")
    t = m.tree
    r_site = parse_column_reduction(root.get("site_reduction"), m.site_count, "site")
    r_node = parse_column_reduction(root.get("node_reduction"), t.node_count, "node")
    r_state = parse_column_reduction(root.get("state_reduction"), m.n, "state")
    ax_node = Axis("node", t.node_count, r_node)
    ax_state = Axis("state", m.n, r_state)
    acc = NdAccum([Axis("site", m.site_count, r_site), ax_node, ax_state], be)
    cs = cross_site(m, be)
    sites = _requested_sites(r_site, m.site_count)
    if sites:
        base = _base_vectors(m, be, sites)
        S = base.shape[0]
        site_L = be.zeros(S)
        cc = {a: be.zeros(S, m.n) for a in range(t.node_count)}
        for c in range(cs.C):
            lh, node, nconst, edge = site_lhood(m, cs, be, base, c, keep_edges=True)
            post = cs.prior[c] * lh
            live = np.array([bool(x != 0) for x in post])   # arbplfmarginal.c:184-191
            site_L = site_L + np.where(live, post, be.num(0))
            fn, fe = site_forward(m, cs, be, base, c, edge)
            for a in range(t.node_count):
                contrib = cs.prior[c] * (fn[a] * node[a])
                contrib = np.where(live[:, None], contrib, be.num(0))
                cc[a] = cc[a] + contrib
        req_n = ax_node.requested()
        req_s = ax_state.requested()
        for k, s in enumerate(sites):
            for a in range(t.node_count):
                if not req_n[a]:
                    continue
                for j in range(m.n):
                    if req_s[j]:
                        acc.accumulate([s, a, j], cc[a][k, j] / site_L[k])
    return acc.to_json()


def _edge_expectations(m, cs, be, base, F, req_csr, trans: bool):
    """_update_site of arbplfdwell.c:201-312 / arbplftrans.c:222-346 -> [S,E] csr order."""
    t = m.tree
    S = base.shape[0]
    site_L = be.zeros(S)
    cc = be.zeros(S, t.edge_count)
    for c in range(cs.C):
        lh, node, nconst, edge = site_lhood(m, cs, be, base, c, keep_edges=True)
        fn, fe = site_forward(m, cs, be, base, c, edge)
        site_L = site_L + cs.prior[c] * lh
        for idx in range(t.edge_count):
            if not req_csr[idx]:
                continue
            b = t.indices[idx]
            fv = _matvec(F[c, idx], node[b])
            x = (fv * fe[idx]).sum(axis=1)
            wgt = cs.prior[c]
            if trans:
                wgt = cs.rates[c] * cs.edge_rates[idx] * cs.prior[c]
            cc[:, idx] = cc[:, idx] + x * wgt
    return cc / site_L[:, None]


def run_dwell(root, mode="mp"):
    be = get_backend(mode)
    m = _top(root, ["site_reduction", "edge_reduction", "state_reduction"], "arbplf-dwell")
    t = m.tree
    n = m.n
    r_site = parse_column_reduction(root.get("site_reduction"), m.site_count, "site")
    r_edge = parse_column_reduction(root.get("edge_reduction"), t.edge_count, "edge")
    r_state = parse_column_reduction(root.get("state_reduction"), n, "state")
    ax_site = Axis("site", m.site_count, r_site)
    ax_edge = Axis("edge", t.edge_count, r_edge)
    ax_state = Axis("state", n, r_state)
    state_agg = r_state.agg_mode != AGG_NONE
    axes = [ax_site, ax_edge] + ([] if state_agg else [ax_state])
    acc = NdAccum(axes, be)
    cs = cross_site(m, be)
    req_user = ax_edge.requested()
    req_csr = [False] * t.edge_count
    for e in range(t.edge_count):
        req_csr[t.order[e]] = req_user[e]
    sites = _requested_sites(r_site, m.site_count)
    if not sites:
        return acc.to_json()
    base = _base_vectors(m, be, sites)
    inv = t.idx_to_user_edge
    if state_agg:
        w, div = agg_weights(r_state, n, be)
        L = be.zeros(n, n)
        for s in range(n):
            L[s, s] = w[s] / div
        F = frechet_matrices(cs, be, L, req_csr)
        X = _edge_expectations(m, cs, be, base, F, req_csr, trans=False)
        for k, s in enumerate(sites):
            for e in range(t.edge_count):
                if req_user[e]:
                    acc.accumulate([s, e], X[k, t.order[e]])
    else:
        req_s = ax_state.requested()
        for st in range(n):
            if not req_s[st]:
                continue
            L = be.zeros(n, n)
            L[st, st] = be.num(1)
            F = frechet_matrices(cs, be, L, req_csr)
            X = _edge_expectations(m, cs, be, base, F, req_csr, trans=False)
            for k, s in enumerate(sites):
                for e in range(t.edge_count):
                    if req_user[e]:
                        acc.accumulate([s, e, st], X[k, t.order[e]])
    return acc.to_json()


def run_trans(root, mode="mp"):
    be = get_backend(mode)
    m = _top(root, ["site_reduction", "edge_reduction", "trans_reduction"], "arbplf-trans")
    t = m.tree
    n = m.n
    r_site = parse_column_reduction(root.get("site_reduction"), m.site_count, "site")
    r_edge = parse_column_reduction(root.get("edge_reduction"), t.edge_count, "edge")
    r_trans = parse_pair_reduction(root.get("trans_reduction"), n, "trans")
    ax_site = Axis("site", m.site_count, r_site)
    ax_edge = Axis("edge", t.edge_count, r_edge)
    ax_trans = Axis("trans", len(r_trans.selection), r_trans,
                    component_names=["first_state", "second_state"],
                    component_indices=[r_trans.first_idx, r_trans.second_idx])
    trans_agg = r_trans.agg_mode != AGG_NONE
    axes = [ax_site, ax_edge] + ([] if trans_agg else [ax_trans])
    acc = NdAccum(axes, be)
    cs = cross_site(m, be)
    req_user = ax_edge.requested()
    req_csr = [False] * t.edge_count
    for e in range(t.edge_count):
        req_csr[t.order[e]] = req_user[e]
    sites = _requested_sites(r_site, m.site_count)
    if not sites:
        return acc.to_json()
    base = _base_vectors(m, be, sites)
    if trans_agg:
        w, div = agg_weights(r_trans, len(r_trans.selection), be)
        L = be.zeros(n, n)
        for k in range(len(r_trans.selection)):
            a, b = r_trans.first_idx[k], r_trans.second_idx[k]
            L[a, b] = L[a, b] + w[k]
        L = L * cs.Q
        L = L / div
        F = frechet_matrices(cs, be, L, req_csr)
        X = _edge_expectations(m, cs, be, base, F, req_csr, trans=True)
        for k, s in enumerate(sites):
            for e in range(t.edge_count):
                if req_user[e]:
                    acc.accumulate([s, e], X[k, t.order[e]])
    else:
        for k2 in range(len(r_trans.selection)):
            a, b = r_trans.first_idx[k2], r_trans.second_idx[k2]
            L = be.zeros(n, n)
            L[a, b] = cs.Q[a, b]
            F = frechet_matrices(cs, be, L, req_csr)
            X = _edge_expectations(m, cs, be, base, F, req_csr, trans=True)
            for k, s in enumerate(sites):
                for e in range(t.edge_count):
                    if req_user[e]:
                        acc.accumulate([s, e, k2], X[k, t.order[e]])
    return acc.to_json()


def per_site_marginal(m: Model, be, sites: Sequence[int]):
    """[S', N, n] posterior marginals (arbplfmarginal.c:142-257), no reductions."""
    t = m.tree
    cs = cross_site(m, be)
    base = _base_vectors(m, be, sites)
    S = base.shape[0]
    site_L = be.zeros(S)
    out = be.zeros(S, t.node_count, m.n)
    for c in range(cs.C):
        lh, node, nconst, edge = site_lhood(m, cs, be, base, c, keep_edges=True)
        post = cs.prior[c] * lh
        live = np.array([bool(x != 0) for x in post])
        site_L = site_L + np.where(live, post, be.num(0))
        fn, fe = site_forward(m, cs, be, base, c, edge)
        for a in range(t.node_count):
            contrib = cs.prior[c] * (fn[a] * node[a])
            out[:, a, :] = out[:, a, :] + np.where(live[:, None], contrib, be.num(0))
    return out / site_L[:, None, None]


def per_site_edge_expect(m: Model, be, sites: Sequence[int], L, trans: bool, requested_csr=None):
    """[S', E] (csr order) dwell / trans expectations for a Frechet direction L."""
    t = m.tree
    cs = cross_site(m, be)
    if requested_csr is None:
        requested_csr = [True] * t.edge_count
    base = _base_vectors(m, be, sites)
    F = frechet_matrices(cs, be, be.asarray(L) if not isinstance(L, np.ndarray) or L.dtype != object else L,
                         requested_csr)
    return _edge_expectations(m, cs, be, base, F, requested_csr, trans=trans), cs


def run_em_update(root, mode="mp"):
    """
    One EM update of the edge rate coefficients (arbplfem.c): per edge, the conditionally expected number of
    transitions over the conditionally expected exit-rate-weighted dwell time, times the current coefficient.
    Frechet directions arbplfem.c:116-128 (L_dwell = diag(-Q_ii), L_trans = offdiag(Q)), accumulation
    :232-379 (every category weighted by prior_c * rate_c, categories of exactly zero likelihood skipped,
    sites weighted by w_s / divisor / site_L), final ratio :434-458, output in user edge order :475-490.
    """
    be = get_backend(mode)
    m = _top(root, ["site_reduction"], "arbplf-em-update")
    t = m.tree
    n = m.n
    E = t.edge_count
    r_site = parse_column_reduction(root.get("site_reduction"), m.site_count, "site")
    if r_site.agg_mode == AGG_NONE:
        raise OracleError("aggregation over sites is required")       # arbplfem.c:566-571
    cs = cross_site(m, be)
    L_dwell = be.zeros(n, n)
    L_trans = be.zeros(n, n)
    for a in range(n):
        L_dwell[a, a] = -cs.Q[a, a]
        for b in range(n):
            if a != b:
                L_trans[a, b] = cs.Q[a, b]
    req = [True] * E
    Fd = frechet_matrices(cs, be, L_dwell, req)
    Ft = frechet_matrices(cs, be, L_trans, req)
    w, div = agg_weights(r_site, m.site_count, be)
    dwell_accum = be.zeros(E)
    trans_accum = be.zeros(E)
    sites = _requested_sites(r_site, m.site_count)
    if sites:
        base = _base_vectors(m, be, sites)
        S = base.shape[0]
        site_L = be.zeros(S)
        dwell_site = be.zeros(S, E)
        trans_site = be.zeros(S, E)
        for c in range(cs.C):
            lh, node, nconst, edge = site_lhood(m, cs, be, base, c, keep_edges=True)
            cat_L = cs.prior[c] * lh
            site_L = site_L + cat_L
            live = np.array([bool(x != 0) for x in cat_L])             # arbplfem.c:322
            fn, fe = site_forward(m, cs, be, base, c, edge)
            wgt = cs.prior[c] * cs.rates[c]                              # arbplfem.c:351
            for idx in range(E):
                b = t.indices[idx]
                xd = (_matvec(Fd[c, idx], node[b]) * fe[idx]).sum(axis=1)
                xt = (_matvec(Ft[c, idx], node[b]) * fe[idx]).sum(axis=1)
                dwell_site[:, idx] = dwell_site[:, idx] + np.where(live, xd * wgt, be.num(0))
                trans_site[:, idx] = trans_site[:, idx] + np.where(live, xt * wgt, be.num(0))
        for k, s in enumerate(sites):
            if not (site_L[k] > 0):
                raise OracleError("site %d has zero likelihood" % s)
            f = w[s] / div / site_L[k]
            dwell_accum = dwell_accum + dwell_site[k] * f
            trans_accum = trans_accum + trans_site[k] * f
    rows = []
    for e in range(E):
        idx = t.order[e]
        v = be.num(0)
        if trans_accum[idx] != 0:
            v = trans_accum[idx] / dwell_accum[idx]
        v = v * cs.edge_rates[idx]
        d = be.to_float(v)
        if d == 0.0:
            d = 0.0
        rows.append([e, d])
    return {"columns": ["edge", "value"], "data": rows}



# --------------------------------------------------------------------------
# second order programs: hess, inv-hess, newton-delta, newton-update
# (arbplfhess.c)
# --------------------------------------------------------------------------

def _site_lhood_subst(m: Model, cs: CrossSite, be, base, cat: int, subst: Dict[int, int]):
    """
    The site likelihood with the transition matrix of edge idx replaced by
    Q^k P_idx for every (idx: k) in subst.  The likelihood is multilinear in
    the per-edge matrices, so this is what evaluate_site_derivatives of
    arbplfhess.c:343-443 computes through its "indirect plane": one
    substitution for a first derivative, two (or Q^2 on one edge) for a second
    one (arbplfhess.c:683-716).
    """
    t = m.tree
    node = {}
    for a in reversed(t.preorder):
        v = base[:, a, :].copy()
        for idx in range(t.indptr[a], t.indptr[a + 1]):
            em = _matvec(cs.P[cat, idx], node[t.indices[idx]])
            for _ in range(subst.get(idx, 0)):
                em = _matvec(cs.Q, em)
            v = v * em
        node[a] = v
    rootv = node[t.root]
    if m.root_mode == ROOT_NONE:
        return rootv.sum(axis=1)
    return rootv @ root_prior_vector(m, cs, be)


def second_order(m: Model, be, r_site: Reduction, want_h=True):
    """
    _recompute_second_order (arbplfhess.c:502-829): log likelihood, its
    gradient and its Hessian with respect to the edge rate coefficients,
    aggregated over the selected sites.  Everything in csr edge order.
    Returns (x, ll, grad[E], hess[E, E]).
    """
    cs = cross_site(m, be)
    t = m.tree
    E = t.edge_count
    w, div = agg_weights(r_site, m.site_count, be)
    sites = _requested_sites(r_site, m.site_count)
    ll = be.num(0)
    grad = be.zeros(E)
    hess = be.zeros(E, E)
    x = [cs.edge_rates[i] for i in range(E)]
    if not sites:
        return x, ll, grad, hess
    base = _base_vectors(m, be, sites)
    S = base.shape[0]
    site_l = be.zeros(S)
    site_g = be.zeros(S, E)
    site_h = be.zeros(S, E, E)
    for c in range(cs.C):
        lh = _site_lhood_subst(m, cs, be, base, c, {}) * cs.prior[c]
        rate = cs.rates[c]
        for k in range(S):
            if lh[k] == 0 and rate != 0:
                # arbplfhess.c:640-660
                raise OracleError("error: infeasible")
        if rate == 0:
            # a zero-rate category has zero derivatives; its likelihood still counts
            site_l = site_l + lh
            continue
        site_l = site_l + lh
        for i in range(E):
            gi = _site_lhood_subst(m, cs, be, base, c, {i: 1})
            site_g[:, i] = site_g[:, i] + cs.prior[c] * rate * gi
            if want_h:
                for j in range(i + 1):
                    sub = {i: 2} if i == j else {i: 1, j: 1}
                    hij = _site_lhood_subst(m, cs, be, base, c, sub)
                    site_h[:, i, j] = site_h[:, i, j] + cs.prior[c] * rate * rate * hij
    for k, s_ in enumerate(sites):
        L = site_l[k]
        ws = w[s_] / div
        ll = ll + ws * be.log(L)
        for i in range(E):
            grad[i] = grad[i] + ws * site_g[k, i] / L
            if want_h:
                for j in range(i + 1):
                    # _lhood_hess_to_ll_hess, arbplfhess.c:455-495
                    v = (site_h[k, i, j] - site_g[k, i] * site_g[k, j] / L) / L
                    hess[i, j] = hess[i, j] + ws * v
    for i in range(E):
        for j in range(i):
            hess[j, i] = hess[i, j]
    return x, ll, grad, hess


def _parse_second_order(root, what):
    """_parse_second_order, arbplfhess.c:1162-1207: the site reduction is required and must aggregate."""
    if not isinstance(root, dict):
        raise OracleError("%s: expected an object" % what)
    _strict_keys(root, ["model_and_data", "site_reduction"], [], what)
    m = parse_model(root["model_and_data"])
    r_site = parse_column_reduction(root["site_reduction"], m.site_count, "site")
    if r_site.agg_mode == AGG_NONE:
        raise OracleError("error: aggregation over sites is required")
    return m, r_site


def _edge_pair_table(m: Model, be, M):
    t = m.tree
    E = t.edge_count
    data = []
    for first in range(E):
        for second in range(E):
            i, j = t.order[first], t.order[second]
            v = M[i, j] if j < i else M[j, i]
            data.append([first, second, be.to_float(v)])
    return {"columns": ["first_edge", "second_edge", "value"], "data": data}


def _edge_table(m: Model, be, v):
    t = m.tree
    return {"columns": ["edge", "value"], "data": [[i, be.to_float(v[t.order[i]])] for i in range(t.edge_count)]}


def _mat_inverse(be, A):
    n = A.shape[0]
    out = be.zeros(n, n)
    for k in range(n):
        e = [be.num(1 if i == k else 0) for i in range(n)]
        col = be.solve(A, e)
        for i in range(n):
            out[i, k] = col[i]
    return out


def run_hess(root, mode="mp"):
    """hess_query, arbplfhess.c:1279-1343."""
    be = get_backend(mode)
    m, r_site = _parse_second_order(root, "arbplf-hess")
    _, _, _, H = second_order(m, be, r_site)
    return _edge_pair_table(m, be, H)


def run_inv_hess(root, mode="mp"):
    """inv_hess_query, arbplfhess.c:1208-1277."""
    be = get_backend(mode)
    m, r_site = _parse_second_order(root, "arbplf-inv-hess")
    _, _, _, H = second_order(m, be, r_site)
    return _edge_pair_table(m, be, _mat_inverse(be, H))


def newton_delta(be, g, H):
    """so_get_newton_delta, arbplfhess.c:203-234: -inv(hess) grad."""
    u = be.solve(H, list(g))
    return [-x for x in u]


def run_newton_delta(root, mode="mp"):
    """newton_delta_query, arbplfhess.c:1345-1398."""
    be = get_backend(mode)
    m, r_site = _parse_second_order(root, "arbplf-newton-delta")
    _, _, g, H = second_order(m, be, r_site)
    return _edge_table(m, be, newton_delta(be, g, H))


def run_newton_update(root, mode="mp"):
    """newton_point_query, arbplfhess.c:1400-1452: x + delta."""
    be = get_backend(mode)
    m, r_site = _parse_second_order(root, "arbplf-newton-update")
    x, _, g, H = second_order(m, be, r_site)
    d = newton_delta(be, g, H)
    return _edge_table(m, be, [x[i] + d[i] for i in range(len(d))])


def compress_patterns(codes):
    """
    Site-pattern compression as the reference's input generators do it
    (examples/BEAST.GTRG/mknuc.py:57-66: an insertion-ordered dictionary
    column -> count): patterns in order of first occurrence, their counts, and
    the pattern index of every site.  Returns (patterns[P,N], counts[P], site_to_pattern[S]).
    """
    codes = np.asarray(codes)
    seen = {}
    counts = []
    smap = np.empty(codes.shape[0], dtype=np.int64)
    for s_, row in enumerate(map(bytes, np.ascontiguousarray(codes))):
        k = seen.get(row)
        if k is None:
            k = seen[row] = len(counts)
            counts.append(0)
        counts[k] += 1
        smap[s_] = k
    first = np.full(len(counts), -1, dtype=np.int64)
    for s_ in range(codes.shape[0] - 1, -1, -1):
        first[smap[s_]] = s_
    return codes[first], np.array(counts, dtype=np.int64), smap


PROGRAMS = {
    "ll": run_ll,
    "deriv": run_deriv,
    "marginal": run_marginal,
    "dwell": run_dwell,
    "trans": run_trans,
    "em_update": run_em_update,
    "hess": run_hess,
    "inv_hess": run_inv_hess,
    "newton_delta": run_newton_delta,
    "newton_update": run_newton_update,
}


def run(program: str, json_in, mode: str = "mp") -> Dict[str, Any]:
    """Run one arbplf program on a JSON document (str or parsed)."""
    if isinstance(json_in, (str, bytes)):
        try:
            json_in = json.loads(json_in)
        except ValueError as e:
            raise OracleError("error on json: %s" % e)
    return PROGRAMS[program](json_in, mode=mode)
