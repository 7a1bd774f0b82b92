"""cfg5 timing: dwell / trans expectations on the cfg2 problem (GTR+G4, 64 taxa x 1M sites), site-summed."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from phyly_b200 import engine as E

class A: pass
args = A(); args.taxa = 64; args.sites = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pb = bench.build_problem(args, 0, 0)
eng = pb["eng"]; defs = np.array(bench.DEFS, dtype=np.float64)
eng.set_data_ptr(defs, pb["codes_t"].data_ptr(), pb["S"], 1)
n = 4
Ld = np.eye(n)                      # dwell in any state: expectation = 1 per edge and site
Lt = np.ones((n, n)) - np.eye(n)    # all transitions (the engine scales by Q itself where the reference does)
for kind, L, name in ((E.KIND_DWELL, Ld, "dwell"), (E.KIND_TRANS, Lt, "trans")):
    for it in range(3):
        t0 = time.perf_counter(); so, tot = eng.edge_expect(kind, L, per_site=False); t1 = time.perf_counter()
        mat, sites = eng.last_timing()
        print("%s: wall %.2f ms, matrices (expm + Frechet + tables) %.3f ms, site kernels %.3f ms, fused kernel %.3f ms, sum[0]=%.6f"
              % (name, (t1 - t0) * 1e3, mat, sites, eng.last_kernel_ms(), tot[0]))
