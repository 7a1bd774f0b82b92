"""Where the end-to-end step spends its time: pipelined loop (waits + transposes + chunk kernels) vs the rest."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
class A: pass
args = A(); args.taxa = 64; args.sites = 1000000
pb = bench.build_problem(args, 0, 0)
eng = pb["eng"]; defs = np.array(bench.DEFS, dtype=np.float64); S = pb["S"]
eng.set_data_ptr(defs, pb["codes_t"].data_ptr(), S, 1); eng.set_site_weights(pb["w_t"].numpy())
for _ in range(3):
    eng.set_edge_rates(pb["edge_rates"]); eng.deriv(per_site=False)
print("resident kernel ms", eng.last_kernel_ms())
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.set_data_async_ptr(defs, pb["codes_t"].data_ptr(), S, pb["w_t"].data_ptr(), 1)
    t1 = time.perf_counter()
    eng.set_edge_rates(pb["edge_rates"]); r = eng.deriv(per_site=False)
    t2 = time.perf_counter()
    mat, sites = eng.last_timing()
    print("e2e step: set_data_async %.3f ms host, deriv call %.3f ms, total %.3f ms | device: matrices %.3f, site part %.3f, pipelined loop %.3f"
          % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, mat, sites, eng.last_kernel_ms()))
