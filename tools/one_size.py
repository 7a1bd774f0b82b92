import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
class A: pass
args = A(); args.taxa = int(sys.argv[1]); args.sites = int(sys.argv[2])
pb = bench.build_problem(args, 0, 0); eng = pb["eng"]
eng.set_data_ptr(np.array(bench.DEFS, dtype=np.float64), pb["codes_t"].data_ptr(), pb["S"], 1)
for it in range(3):
    eng.set_edge_rates(pb["edge_rates"]); r = eng.deriv(per_site=False)
print("deriv kernel %.3f ms" % eng.last_kernel_ms())
