"""
N > 1 host logic on CPU: two processes over gloo, each owning a block of site patterns, summing
their (1 + E) partial results with one all-reduce -- the same partition / collective the GPU path
uses (there the partial sums come from the CUDA engine and the collective is
ncclAllReduce issued by plf_engine.cu).  The per-shard evaluation here is the C restatement.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(S, world, rank):
    per = (S + world - 1) // world
    return rank * per, min(S, (rank + 1) * per)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_port
    from tests import helpers as H
    from oracle import arbplf_oracle as O
    prob = H.random_problem(12, ntips=24, n=4, S=70, ncat=2, root="equilibrium_distribution", missing=0.3)
    m = O.parse_model(prob["model_and_data"])
    ref = H.seam_reference("random12", prob)
    params, cs = H.model_params(m)
    defs, codes = H.dedupe_rows(m.dense_pmat())
    S = m.site_count
    w = np.linspace(0.5, 2.0, S)
    lo, hi = _shard(S, world, rank)
    t = m.tree
    _, sum_ll, sum_d = c_port.ll_deriv(t.indptr, t.indices, t.preorder, ref["P"], ref["Dm"], params["cat_prior"],
                                       m.root_mode, params["root_vec"], codes[lo:hi], defs, w=w[lo:hi], nthreads=1)
    buf = torch.tensor(np.concatenate([[sum_ll], sum_d]), dtype=torch.float64)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        want_ll = float((w * ref["ll"]).sum())
        want_d = (w[:, None] * ref["D"]).sum(axis=0)
        got = buf.numpy()
        ok = abs(got[0] - want_ll) <= 1e-11 * abs(want_ll) and np.all(
            np.abs(got[1:] - want_d) <= 1e-11 * np.abs(want_d) + 2e-14 * (w[:, None] * ref["Dabs"]).sum(axis=0))
        out.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_site_sharding_with_one_allreduce(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shards_cover_all_sites_once():
    for S in (1, 7, 70, 1000001):
        for world in (1, 2, 3, 8):
            seen = 0
            prev = 0
            for r in range(world):
                lo, hi = _shard(S, world, r)
                lo, hi = min(lo, S), max(min(hi, S), min(lo, S))
                assert lo == prev or lo >= S
                seen += hi - lo
                prev = hi if hi > lo else prev
            assert seen == S
