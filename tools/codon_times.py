"""cfg4 timing: 61-state codon model, 256 taxa x 100k sites, ll (and ll+deriv) through the device seam."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import phyly_b200.arbplf as A
from phyly_b200.engine import Engine
from tests.test_scale_gpu import _codon_model

taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
Q, pi = _codon_model()
n = Q.shape[0]
edges, N = bench.yule_tree(taxa, seed=21)
rng = np.random.default_rng(22)
defs = np.vstack([np.eye(n), np.ones((1, n))])
md = {"edges": edges, "edge_rate_coefficients": [float(x) for x in rng.exponential(0.05, len(edges))],
      "rate_matrix": Q.tolist(), "root_prior": "equilibrium_distribution", "rate_divisor": "equilibrium_exit_rate",
      "character_definitions": defs.tolist(), "character_data": [[n] * N]}
s = json.loads(A.arbplf_model_summary(json.dumps({"model_and_data": md})))
eng = Engine(0)
eng.set_tree(s["indptr"], s["indices"], s["preorder"])
eng.set_model(np.array(s["q_hi"]).reshape(n, n), np.array(s["q_lo"]).reshape(n, n), s["edge_rates_csr"], s["cat_rates"],
              s["cat_prior"], s["root_mode"], s["root_vec"])
codes = np.full((S, N), n, dtype=np.uint8)
leaves = [a for a in range(N) if s["indptr"][a] == s["indptr"][a + 1]]
for a in leaves:
    codes[:, a] = rng.integers(0, n, S)
eng.set_data(defs, codes)
E = N - 1
for it in range(3):
    t0 = time.perf_counter()
    _, tot = eng.ll(per_site=False)
    wall = time.perf_counter() - t0
    ms_mat, ms_sites = eng.last_timing()
flops = float(S) * E * 2 * n * n
print(json.dumps({"what": "cfg4 ll", "taxa": taxa, "sites": S, "ms_matrices": ms_mat, "ms_sites": ms_sites, "wall_ms": wall * 1e3,
                  "updates_per_s": S * E / (ms_sites * 1e-3), "tflops_dense": flops / (ms_sites * 1e-3) / 1e12, "sum_ll": tot}))
if len(sys.argv) > 3:
    for it in range(2):
        r = eng.deriv(per_site=False)
        ms_mat, ms_sites = eng.last_timing()
    print(json.dumps({"what": "cfg4 ll+deriv", "ms_matrices": ms_mat, "ms_sites": ms_sites,
                      "updates_per_s": S * E / (ms_sites * 1e-3), "sum_ll": r["sum_ll"]}))
