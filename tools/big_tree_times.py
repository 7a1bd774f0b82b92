"""Fused 4-state kernels on trees too large for the constant-memory / fully staged configurations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
for taxa, sites in ((64, 1000000), (128, 500000), (256, 250000), (512, 125000)):
    class A: pass
    args = A(); args.taxa = taxa; args.sites = sites
    pb = bench.build_problem(args, 0, 0)
    eng = pb["eng"]; defs = np.array(bench.DEFS, dtype=np.float64)
    eng.set_data_ptr(defs, pb["codes_t"].data_ptr(), pb["S"], 1)
    for it in range(3):
        eng.set_edge_rates(pb["edge_rates"]); r = eng.deriv(per_site=False); kd = eng.last_kernel_ms()
        s, t = eng.ll(per_site=False); kl = eng.last_kernel_ms()
    upd = float(sites) * pb["E"] * pb["C"]
    print("taxa %4d sites %8d: deriv kernel %.3f ms (%.2e upd/s)  ll kernel %.3f ms (%.2e upd/s)" % (taxa, sites, kd, upd / kd * 1e3, kl, upd / kl * 1e3))
    eng.close()
