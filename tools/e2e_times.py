"""Where the end-to-end step (upload + matrices + ll+deriv + read-back) spends its time at the per-rank sizes of the
strong-scaling run: host time inside each C-ABI call, for S = 10^6 / 8, / 4, / 2, / 1.
Run on the GPU box: python tools/e2e_times.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench


class A:
    pass


args = A(); args.taxa = 64; args.sites = 1000000
pb = bench.build_problem(args, 0, 0)
eng = pb["eng"]; N = pb["N"]
defs = np.array(bench.DEFS, dtype=np.float64)
for S in (125000, 250000, 500000, 1000000):
    cp = pb["codes_t"].data_ptr(); wp = pb["w_t"].data_ptr()
    eng.set_data_ptr(defs, cp, S, 1); eng.set_site_weights(pb["w_t"].numpy()[:S])
    for _ in range(5):
        eng.set_edge_rates(pb["edge_rates"]); eng.deriv(per_site=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30):
        eng.set_edge_rates(pb["edge_rates"]); eng.deriv(per_site=False)
    torch.cuda.synchronize()
    resident = (time.perf_counter() - t0) / 30
    for _ in range(5):
        eng.set_data_async_ptr(defs, cp, S, wp, 1); eng.set_edge_rates(pb["edge_rates"]); eng.deriv(per_site=False)
    ta = tb = tc = 0.0
    for _ in range(30):
        t0 = time.perf_counter()
        eng.set_data_async_ptr(defs, cp, S, wp, 1)
        t1 = time.perf_counter()
        eng.set_edge_rates(pb["edge_rates"])
        t2 = time.perf_counter()
        eng.deriv(per_site=False)
        t3 = time.perf_counter()
        ta += t1 - t0; tb += t2 - t1; tc += t3 - t2
    h2d = S * (N + 8) / 1e6
    print("S=%7d resident %.3f ms | e2e %.3f ms = set_data_async %.3f + set_edge_rates %.3f + deriv %.3f | H2D %.1f MB"
          % (S, resident * 1e3, (ta + tb + tc) / 30 * 1e3, ta / 30 * 1e3, tb / 30 * 1e3, tc / 30 * 1e3, h2d))
