"""cfg3 timing: HKY85+I marginals, 128 taxa x S sites (site-summed and per-site)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bench
import test_scale_gpu as T
import phyly_b200.arbplf as A
from phyly_b200.engine import Engine

S = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
doc, N = T._hky_doc(128)
s = json.loads(A.arbplf_model_summary(json.dumps(doc)))
eng = Engine(0)
eng.set_tree(s["indptr"], s["indices"], s["preorder"])
eng.set_model(np.array(s["q_hi"]).reshape(4, 4), np.array(s["q_lo"]).reshape(4, 4), s["edge_rates_csr"],
              s["cat_rates"], s["cat_prior"], s["root_mode"], s["root_vec"])
P = eng.transition_matrices()
codes = np.empty((S, N), dtype=np.uint8)
bench.simulate_codes(s, P, S, seed=13, out=codes)
eng.set_data(np.array(bench.DEFS, dtype=np.float64), codes)
for it in range(3):
    t0 = time.perf_counter(); sm, tot = eng.marginal(per_site=False); t1 = time.perf_counter()
    print("marginal (site-summed): wall %.1f ms, device sites-part %.1f ms, kernel %.3f ms" % ((t1 - t0) * 1e3, eng.last_timing()[1], eng.last_kernel_ms()))
t0 = time.perf_counter(); r = eng.deriv(per_site=False); t1 = time.perf_counter()
print("deriv (fused): wall %.1f ms, kernel %.3f ms" % ((t1 - t0) * 1e3, eng.last_kernel_ms()))
Ssub = min(S, 100000)
print("sum check", float(tot.sum()), "expected", S * N)
