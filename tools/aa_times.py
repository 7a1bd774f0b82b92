"""20-state (amino-acid sized) model: ll and ll+deriv through the 32-row DMMA tile kernels."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import phyly_b200.arbplf as A
from phyly_b200.engine import Engine
taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
rng = np.random.default_rng(77)
n = 20
pi = rng.dirichlet(np.ones(n) * 4)
R = rng.random((n, n)) + 0.05; R = (R + R.T) / 2
Q = R * pi[None, :]; np.fill_diagonal(Q, 0.0)
edges, N = bench.yule_tree(taxa, seed=5)
defs = np.vstack([np.eye(n), np.ones((1, n))])
md = {"edges": edges, "edge_rate_coefficients": [float(x) for x in rng.exponential(0.1, len(edges))],
      "rate_matrix": Q.tolist(), "root_prior": "equilibrium_distribution", "rate_divisor": "equilibrium_exit_rate",
      "rate_mixture": {"rates": [0.2, 0.7, 1.2, 1.9], "prior": [0.25, 0.25, 0.25, 0.25]},
      "character_definitions": defs.tolist(), "character_data": [[n] * N]}
s = json.loads(A.arbplf_model_summary(json.dumps({"model_and_data": md})))
eng = Engine(0)
eng.set_tree(s["indptr"], s["indices"], s["preorder"])
eng.set_model(np.array(s["q_hi"]).reshape(n, n), np.array(s["q_lo"]).reshape(n, n), s["edge_rates_csr"], s["cat_rates"],
              s["cat_prior"], s["root_mode"], s["root_vec"])
codes = np.full((S, N), n, dtype=np.uint8)
for a in range(N):
    if s["indptr"][a] == s["indptr"][a + 1]:
        codes[:, a] = rng.integers(0, n, S)
eng.set_data(defs, codes)
E, C = N - 1, 4
for label in ("tile", "scalar"):
    if label == "scalar":
        os.environ["PLF_NO_TILE"] = "1"
    for it in range(2):
        _, tot = eng.ll(per_site=False); ms_ll = eng.last_timing()[1]
        r = eng.deriv(per_site=False); ms_d = eng.last_timing()[1]
    print("%s: ll %.2f ms (%.2e upd/s)  ll+deriv %.2f ms (%.2e upd/s)  sum_ll %.6f" % (label, ms_ll, S * E * C / ms_ll * 1e3, ms_d, S * E * C / ms_d * 1e3, tot))
