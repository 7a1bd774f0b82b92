"""More than 4 rate categories on the fused 4-state path (category windows): the cfg2 workload with GTR+Gamma4+I
(5 categories, the shape of examples/BEAST.GTRGI) and with Gamma8, ll + deriv, data resident; and the forced generic path."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import phyly_b200.arbplf as A
from phyly_b200 import engine as E
from phyly_b200.engine import Engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
for label, mix in (("GTR+G4+I (C=5)", {"gamma_categories": 4, "gamma_shape": 0.5, "invariable_prior": 0.2}),
                   ("GTR+G8 (C=8)", {"gamma_categories": 8, "gamma_shape": 0.5})):
    doc, N = bench.model_document(64)
    doc["model_and_data"]["normalized_median_gamma_rate_mixture"] = mix
    s = json.loads(A.arbplf_model_summary(json.dumps(doc)))
    C = s["category_count"]
    eng = Engine(0)
    eng.set_tree(s["indptr"], s["indices"], s["preorder"])
    eng.set_model(np.array(s["q_hi"]).reshape(4, 4), np.array(s["q_lo"]).reshape(4, 4), s["edge_rates_csr"], s["cat_rates"],
                  s["cat_prior"], s["root_mode"], s["root_vec"])
    P = eng.transition_matrices()
    codes = np.empty((S, N), dtype=np.uint8)
    bench.simulate_codes(s, P, S, seed=3, out=codes)
    eng.set_data(np.array(bench.DEFS, dtype=np.float64), codes)
    out = {"model": label, "sites": S, "categories": C}
    for path, name in ((E.PATH_AUTO, "fused"), (E.PATH_GENERIC, "generic")):
        if name == "generic" and S > 200000:
            continue
        eng.set_path(path)
        for it in range(3):
            eng.set_edge_rates(s["edge_rates_csr"])
            r = eng.deriv(per_site=False)
            ms_mat, ms_sites = eng.last_timing()
        out[name] = {"ms_sites": ms_sites, "updates_per_s": S * (N - 1) * C / (ms_sites * 1e-3), "kernel": eng.last_kernel_name(),
                     "sum_ll": r["sum_ll"]}
    print(json.dumps(out), flush=True)
    eng.close()
