import sys, json, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
class A: pass
args=A(); args.taxa=64; args.sites=1000000
pb = bench.build_problem(args, 0, 0)
eng=pb['eng']; defs=np.array(bench.DEFS,dtype=np.float64)
eng.set_data_ptr(defs, pb['codes_t'].data_ptr(), pb['S'], 1)
for it in range(3):
    eng.set_edge_rates(pb['edge_rates']); r=eng.deriv(per_site=False); kd=eng.last_kernel_ms()
    s,t=eng.ll(per_site=False); kl=eng.last_kernel_ms()
print('deriv kernel ms %.3f  ll kernel ms %.3f  sum_ll %.6f %.6f'%(kd,kl,r['sum_ll'],t))
