"""Fused edge kernel: time every candidate configuration (PLF_F4_CONFIG=i) on one problem size."""
import sys, os, subprocess
taxa, sites = sys.argv[1], sys.argv[2]
code = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, bench
class A: pass
args = A(); args.taxa = int(%s); args.sites = int(%s)
pb = bench.build_problem(args, 0, 0); eng = pb["eng"]
eng.set_data_ptr(np.array(bench.DEFS, dtype=np.float64), pb["codes_t"].data_ptr(), pb["S"], 1)
try:
    for it in range(3):
        eng.set_edge_rates(pb["edge_rates"]); r = eng.deriv(per_site=False)
    print("%%.3f ms" %% eng.last_kernel_ms())
except Exception as ex:
    print("n/a (%%s)" %% str(ex)[:60])
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), taxa, sites)
for i in range(13):
    env = dict(os.environ, PLF_F4_CONFIG=str(i))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    print("config %2d: %s" % (i, out[-1] if out else "?"))
