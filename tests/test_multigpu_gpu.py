"""
Two ranks, two GPUs: each rank owns half of the site patterns; the in-stream all-reduce of plf_engine.cu
(plf_comm_init: the peer-memory kernel, and ncclAllReduce as its fallback) must give every rank the sums of a
single-GPU run over all sites -- for the log-likelihood + derivative query, a dwell query, the site-summed
marginals and the Hessian query -- and the same bits on every rank.  Skipped on a one-GPU box
(the driver's `pytest -m gpu` run); run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py`.
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import bench
from phyly_b200 import engine as E
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
class A: pass
args = A(); args.taxa = 24; args.sites = 40000
pb = bench.build_problem(args, local, 0)             # every rank builds the SAME problem (rank 0's seed)
eng = pb["eng"]; S = pb["S"]; defs = np.array(bench.DEFS, dtype=np.float64)
codes = pb["codes"]
w = 1.0 + np.random.default_rng(3).poisson(2.0, S).astype(np.float64)
def queries(e):
    r = e.deriv(per_site=False)
    _, dw = e.edge_expect(E.KIND_DWELL, np.eye(4), per_site=False)
    _, mg = e.marginal(per_site=False)
    hl, hg, H = e.hess()                                  # 47 x 47 doubles: the peer-memory all-reduce; ll+deriv: its fused form
    return np.concatenate([[r["sum_ll"]], r["sum_deriv"], dw, mg.ravel(), [hl], hg, H.ravel()])
# single-GPU reference over all sites (no communicator yet)
eng.set_data(defs, codes); eng.set_site_weights(w)
want = queries(eng)
# sharded: this rank's half, then the communicator
lo, hi = rank * S // world, (rank + 1) * S // world
uid = [E.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(world, rank, uid[0])
eng.set_data(defs, np.ascontiguousarray(codes[lo:hi])); eng.set_site_weights(w[lo:hi])
got = queries(eng)
# every rank holds the same bits (the peer-memory all-reduce adds the vectors in rank order on every rank)
peers = [None] * world
dist.all_gather_object(peers, got.tobytes())
same = all(p == peers[0] for p in peers)
# the NCCL fallback gives the same sums
os.environ["PLF_NO_PEER_ALLREDUCE"] = "1"
got_nccl = queries(eng)
del os.environ["PLF_NO_PEER_ALLREDUCE"]
scale = np.abs(want).max()
same = same and bool(np.all(np.abs(got_nccl - got) <= 1e-12 * np.abs(got) + 1e-13 * scale))
ok = same and bool(np.all(np.abs(got - want) <= 1e-11 * np.abs(want) + 1e-13 * scale))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_RESULT", int(flag.item()), float(np.abs(got - want).max()))
dist.destroy_process_group()
'''


def test_two_rank_allreduce_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("MULTIGPU_RESULT")]
    assert lines and lines[0].split()[1] == "1", r.stdout[-2000:] + r.stderr[-2000:]
