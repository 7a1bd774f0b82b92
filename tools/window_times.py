import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
for taxa, sites in ((128, 500000), (256, 250000)):
    class A: pass
    args = A(); args.taxa = taxa; args.sites = sites
    pb = bench.build_problem(args, 0, 0)
    eng = pb["eng"]; defs = np.array(bench.DEFS, dtype=np.float64)
    eng.set_data_ptr(defs, pb["codes_t"].data_ptr(), pb["S"], 1)
    for wf in ("", "2", "1"):
        if wf: os.environ["PLF_F4_WINDOW"] = wf
        else: os.environ.pop("PLF_F4_WINDOW", None)
        for it in range(3):
            eng.set_edge_rates(pb["edge_rates"]); r = eng.deriv(per_site=False); ms = eng.last_timing()[1]
        upd = float(sites) * pb["E"] * pb["C"]
        print("taxa %4d sites %8d window %s: deriv sites %.3f ms (%.2e upd/s) %s sum_ll %.6f" % (taxa, sites, wf or "-", ms, upd / ms * 1e3, eng.last_kernel_name(), r["sum_ll"]), flush=True)
    eng.close()
