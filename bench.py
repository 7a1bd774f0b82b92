#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 phylogenetic likelihood engine.

Metric (BASELINE.json): site-edge-category updates/s for ll + deriv.  One update
= one (site pattern, edge, rate category) triple processed once
(evaluate_site_lhood.c:54 in the reference); updates per evaluation = S*E*C.

Workload at N=1 (BASELINE.json configs[1], "cfg2"): GTR+Gamma4 model of
examples/BEAST.GTRG on a synthetic 64-taxon tree x 1,000,000 site patterns,
arbplf-ll + arbplf-deriv with the site axis summed.  A "step" is one full
evaluation: edge rates change -> E*C matrix exponentials (+ Q.P) -> tip tables
-> fused pruning + outside pass over all sites -> reduction (-> ncclAllReduce
when N > 1) -> (1+E) doubles back on the host.

  value : inputs (codes, weights) already resident in HBM.
  e2e   : the same evaluation through the C ABI with HOST buffers: the step
          also copies the alignment codes and the site weights from pinned
          host memory to the device and reads the results back.

With --impl reference the same metric is measured for the reference's CPU
algorithm: the C restatement in oracle/c (the Arb reference itself cannot be
built in this image), all host threads, on a bounded sample of the same
workload.

N > 1: one process per GPU (torchrun), each rank owns S site patterns of its
own (weak scaling), results are summed with one in-stream ncclAllReduce.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# examples/BEAST.GTRG/in.json (model part)
PI = [0.26121198857732775, 0.24655436448599188, 0.33909428971935046, 0.1531393572173299]
RATE_MATRIX = [[0.0 if i == j else PI[j] for j in range(4)] for i in range(4)]
DEFS = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 1, 1]]

# algorithmic HBM bytes per update of the streamed formulation
BYTES_PER_UPDATE_LL_DERIV = 79.4
BYTES_PER_UPDATE_LL = 31.9


def yule_tree(taxa, seed):
    """Random binary tree, edges parent -> child listed in random user order."""
    rng = np.random.default_rng(seed)
    children = {0: []}
    leaves = [0]
    nxt = 1
    while len(leaves) < taxa:
        a = leaves.pop(int(rng.integers(len(leaves))))
        children[a] = [nxt, nxt + 1]
        for k in (nxt, nxt + 1):
            children[k] = []
            leaves.append(k)
        nxt += 2
    N = nxt
    perm = rng.permutation(N)
    edges = [[int(perm[a]), int(perm[b])] for a in children for b in children[a]]
    order = rng.permutation(len(edges))
    return [edges[i] for i in order], N


def model_document(taxa):
    edges, N = yule_tree(taxa, seed=1)
    rng = np.random.default_rng(2)
    rates = rng.exponential(0.05, len(edges))
    md = {
        "edges": edges,
        "edge_rate_coefficients": [float(x) for x in rates],
        "rate_matrix": RATE_MATRIX,
        "root_prior": PI,
        "rate_divisor": "equilibrium_exit_rate",
        "normalized_median_gamma_rate_mixture": {"gamma_categories": 4, "gamma_shape": 0.5},
        "character_definitions": DEFS,
        "character_data": [[4] * N],
    }
    return {"model_and_data": md}, N


def simulate_codes(summary, P, S, seed, out):
    """Simulate S columns down the tree under the model; leaves observed (1% missing), internal nodes unobserved."""
    rng = np.random.default_rng(seed)
    indptr, indices, preorder = summary["indptr"], summary["indices"], summary["preorder"]
    N = len(preorder)
    C = P.shape[0]
    cum = np.cumsum(P, axis=3)
    cat = rng.integers(0, C, S)
    state = np.empty((N, S), dtype=np.int8)
    root = preorder[0]
    state[root] = np.searchsorted(np.cumsum(PI), rng.random(S)).clip(0, 3)
    for a in preorder:
        for idx in range(indptr[a], indptr[a + 1]):
            b = indices[idx]
            rows = cum[cat, idx, state[a].astype(np.int64)]          # [S,4]
            u = rng.random(S)
            state[b] = (u[:, None] > rows).sum(axis=1).clip(0, 3)
    for a in range(N):
        if indptr[a] == indptr[a + 1]:
            col = state[a].astype(np.uint8)
            miss = rng.random(S) < 0.01
            col[miss] = 4
            out[:, a] = col
        else:
            out[:, a] = 4


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []          # (arrival time, text)
        self.windows = []        # timed regions [t0, t1] (time.perf_counter)

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples that arrived inside a timed region; a region shorter than the sampling period may hold none,
        # then every sample taken while the GPU was under load (warm-up to the last timed step) is used
        inside = [ln for t, ln in self.lines if any(a <= t <= b + 0.02 for a, b in self.windows)]
        window = "timed regions"
        if not inside and self.windows:
            lo, hi = self.windows[0][0] - 1.0, self.windows[-1][1] + 0.02
            inside = [ln for t, ln in self.lines if lo <= t <= hi]
            window = "warm-up + timed regions (timed regions shorter than the sampling period)"
        sm, mx, reasons = [], None, set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "window": window, "reasons": sorted(reasons)}


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def dist_setup(n_gpus):
    """torch.distributed over NCCL (one process per GPU), gloo for host objects."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    return rank, world, local


def host_matrices(summary):
    """P and rate*Q*P on the host (scipy), used only to simulate data and by the CPU reference arm."""
    import scipy.linalg
    n, C = summary["state_count"], summary["category_count"]
    Q = np.array(summary["q_hi"]).reshape(n, n)
    t = np.array(summary["edge_rates_csr"])
    r = np.array(summary["cat_rates"])
    P = np.empty((C, len(t), n, n))
    D = np.empty_like(P)
    for c in range(C):
        for e in range(len(t)):
            P[c, e] = scipy.linalg.expm(Q * (r[c] * t[e]))
            D[c, e] = r[c] * (Q @ P[c, e])
    return P, D


def build_problem(args, local, rank, use_engine=True):
    """Host-side setup through the product's own C host layer (no oracle involved)."""
    import phyly_b200.arbplf as A
    doc, N = model_document(args.taxa)
    summary = json.loads(A.arbplf_model_summary(json.dumps(doc)))
    n, C = summary["state_count"], summary["category_count"]
    edge_rates = np.array(summary["edge_rates_csr"])
    S = args.sites
    P, D = host_matrices(summary)
    eng = None
    if use_engine:
        import torch
        from phyly_b200.engine import Engine
        eng = Engine(local)
        eng.set_tree(summary["indptr"], summary["indices"], summary["preorder"])
        q_hi = np.array(summary["q_hi"]).reshape(n, n)
        q_lo = np.array(summary["q_lo"]).reshape(n, n)
        eng.set_model(q_hi, q_lo, edge_rates, summary["cat_rates"], summary["cat_prior"], summary["root_mode"],
                      summary["root_vec"])
        codes_t = torch.empty((S, N), dtype=torch.uint8, pin_memory=True)
        codes = codes_t.numpy()
        w_t = torch.ones(S, dtype=torch.float64, pin_memory=True)
    else:
        codes_t, w_t = None, None
        codes = np.empty((S, N), dtype=np.uint8)
    simulate_codes(summary, P, S, seed=3 + rank, out=codes)
    return dict(eng=eng, summary=summary, N=N, E=N - 1, n=n, C=C, S=S, codes_t=codes_t, codes=codes, w_t=w_t,
                edge_rates=edge_rates, P=P, D=D)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from phyly_b200 import engine as E
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    pb = build_problem(args, local, rank)
    eng = pb["eng"]
    S, Eg, C, N = pb["S"], pb["E"], pb["C"], pb["N"]
    defs = np.array(DEFS, dtype=np.float64)
    if world > 1:
        uid = [E.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(world, rank, uid[0])
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        eng.set_edge_rates(pb["edge_rates"])          # invalidates P, Q.P and the tip tables
        return eng.deriv(per_site=False)

    sync_upload = bool(os.environ.get("PLF_BENCH_SYNC_UPLOAD"))

    def step_e2e():
        if sync_upload:
            eng.set_data_ptr(defs, pb["codes_t"].data_ptr(), S, 1)
            eng.set_site_weights(pb["w_t"].numpy())
        else:       # chunked upload overlapped with the kernel
            eng.set_data_async_ptr(defs, pb["codes_t"].data_ptr(), S, pb["w_t"].data_ptr(), 1)
        eng.set_edge_rates(pb["edge_rates"])
        return eng.deriv(per_site=False)

    # ---- device-resident measurement ----
    eng.set_data_ptr(defs, pb["codes_t"].data_ptr(), S, 1)
    eng.set_site_weights(pb["w_t"].numpy())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        res = step_resident()
    barrier()
    eng.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, mat_ms = [], []
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        res = step_resident()
        kern_ms.append(eng.last_kernel_ms())
        mat_ms.append(eng.last_timing()[0])
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark(t0, t0 + wall)
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count()
    # ---- end-to-end measurement (host buffers) ----
    e2e_steps = max(1, min(args.steps, 50))
    for _ in range(min(args.warmup, 3)):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t1 = time.perf_counter()
    f0.record(stream)
    for _ in range(e2e_steps):
        res2 = step_e2e()
    f1.record(stream)
    barrier()
    sampler.mark(t1, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    # ---- supplementary: log-likelihood only (the metric's other half), device-resident, same timing rules ----
    ll_steps = max(1, min(args.steps, 50))

    def step_ll():
        eng.set_edge_rates(pb["edge_rates"])
        return eng.ll(per_site=False)

    for _ in range(3):
        step_ll()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ll_kern = []
    g0.record(stream)
    for _ in range(ll_steps):
        step_ll()
        ll_kern.append(eng.last_kernel_ms())
    g1.record(stream)
    barrier()
    ms_ll = g0.elapsed_time(g1)
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        tt = torch.tensor([ms, ms_e2e, ms_ll], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_ll = float(tt[0]), float(tt[1]), float(tt[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    updates_per_step = float(S) * Eg * C * world
    value = updates_per_step * args.steps / (ms * 1e-3)
    e2e_value = updates_per_step * e2e_steps / (ms_e2e * 1e-3)
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    hbm_peak = peaks["hbm_gbs"] if peaks and "hbm_gbs" in peaks else 6650.0
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    fp64 = load_json(os.path.join(ROOT, "profiles", "fp64_peak.json")) or {}
    traffic = load_json(os.path.join(ROOT, "profiles", "ncu_traffic.json")) or {}
    k_ms = float(np.mean(kern_ms)) if kern_ms else None
    alg_bytes = BYTES_PER_UPDATE_LL_DERIV * float(S) * Eg * C
    # fp64 work of the fused kernel per (site, category): see DESIGN.md section 5
    n_int_edges = sum(1 for b in pb["summary"]["indices"]
                      if pb["summary"]["indptr"][b] != pb["summary"]["indptr"][b + 1])
    n_tip_edges = Eg - n_int_edges
    flops = float(S) * C * (n_int_edges * (36 + 108) + n_tip_edges * (4 + 14))
    roofline = {
        "bound": "hbm", "kernel": traffic.get("kernel", "fused4_kernel"),
        "achieved": alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None, "peak": hbm_peak, "unit": "GB/s",
        "frac": (alg_bytes / (k_ms * 1e-3) / 1e9) / hbm_peak if k_ms else None,
        "traffic": traffic.get("dram_bytes_per_launch"),
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes,
        "note": ("algorithmic bytes are SURVEY 8(d)'s streamed-partials figure (79.4 B/update); the fused kernel keeps "
                 "partials on chip, so a fraction above 1 means it moves less than that formulation (traffic = DRAM "
                 "bytes per launch measured by ncu); ncu shows it bound by slab-load latency and LSU wavefronts, "
                 "fp64 gives its flop rate against the measured DFMA peak"),
        "measured_dram_gbs": (traffic["dram_bytes_per_launch"] / (k_ms * 1e-3) / 1e9
                              if (k_ms and traffic.get("dram_bytes_per_launch")) else None),
        "measured_dram_frac": (traffic["dram_bytes_per_launch"] / (k_ms * 1e-3) / 1e9 / hbm_peak
                               if (k_ms and traffic.get("dram_bytes_per_launch")) else None),
        "fp64": {"achieved_tflops": flops / (k_ms * 1e-3) / 1e12 if k_ms else None,
                 "peak_tflops": fp64.get("dfma_tflops"),
                 "frac": (flops / (k_ms * 1e-3) / 1e12) / fp64["dfma_tflops"] if (k_ms and fp64.get("dfma_tflops")) else None,
                 "flops_per_launch": flops},
        "kernel_ms": k_ms, "matrix_kernels_ms": float(np.mean(mat_ms)) if mat_ms else None,
    }
    line = {
        "metric": "site-edge-category updates/s (ll+deriv)", "value": value, "unit": "updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: GTR+Gamma4 (examples/BEAST.GTRG model) x %d-taxon Yule tree x %d site patterns per GPU, "
                               "arbplf-ll + arbplf-deriv, site axis summed" % (args.taxa, S),
                   "sites_per_gpu": S, "taxa": args.taxa, "edges": Eg, "categories": C, "states": pb["n"],
                   "sharding": "sites across GPUs, one ncclAllReduce of %d doubles" % (1 + Eg),
                   "cache": "inputs and scratch (%.0f MB) exceed the 126 MB L2; no explicit flush" % (S * N / 1e6 + 600)},
        "e2e": {"value": e2e_value, "unit": "updates/s", "h2d_bytes_per_step": int(S * N + 8 * S + 8 * Eg),
                "d2h_bytes_per_step": int(8 * (1 + Eg)), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "api": ("plf_set_data + plf_set_site_weights" if sync_upload else "plf_set_data_async") + " + plf_set_edge_rates + plf_deriv (include/plf.h), pinned host buffers"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "ll_only": {"metric": "site-edge-category updates/s (ll)", "value": updates_per_step * ll_steps / (ms_ll * 1e-3),
                    "unit": "updates/s", "steps": ll_steps, "ms_per_step": ms_ll / ll_steps,
                    "kernel_ms": float(np.mean(ll_kern)),
                    "roofline_frac_hbm": (BYTES_PER_UPDATE_LL * float(S) * Eg * C / (float(np.mean(ll_kern)) * 1e-3) / 1e9) / hbm_peak,
                    "note": "supplementary; algorithmic bytes 31.9 B/update (SURVEY 8d); no slab, partials never leave the SM"},
        "result_check": {"sum_ll": res["sum_ll"], "sum_ll_e2e": res2["sum_ll"], "wall_s": wall},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(pb, target_seconds=args.cpu_seconds)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_port_problem(pb):
    """Inputs of the C restatement taken from the engine (matrices) and the workload (codes)."""
    return dict(P=pb["P"], D=pb["D"], prior=np.array(pb["summary"]["cat_prior"]),
                root_mode=pb["summary"]["root_mode"], root_vec=np.array(pb["summary"]["root_vec"]))


def host_threads():
    """All host cores this process may use.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to
    its workers, which would silently turn the reference arm into a one-thread run for N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(pb, target_seconds=15.0, sample=None):
    """Time oracle/c (the 'port' of the reference algorithm) on a bounded sample, all host threads."""
    from oracle import c_port
    cp = cpu_port_problem(pb)
    s = pb["summary"]
    threads = host_threads()
    defs = np.array(DEFS, dtype=np.float64)

    def run(nsites):
        t0 = time.perf_counter()
        c_port.ll_deriv(s["indptr"], s["indices"], s["preorder"], cp["P"], cp["D"], cp["prior"], cp["root_mode"],
                        cp["root_vec"], pb["codes"][:nsites], defs, w=None, nthreads=threads, want_site_ll=False)
        return time.perf_counter() - t0
    if sample is None:
        pilot = min(pb["S"], 4000 * threads)
        run(min(pilot, 1000))
        dt = run(pilot)
        sample = int(min(pb["S"], max(pilot, pilot * target_seconds / max(dt, 1e-6))))
    dt = run(sample)
    return {"value": sample * pb["E"] * pb["C"] / dt, "unit": "updates/s", "cores": threads, "kind": "port",
            "sample": "%d of %d site patterns, ll+deriv, oracle/c/plf_oracle.c (OpenMP, fp64), %.2f s" % (sample, pb["S"], dt),
            "seconds": dt, "sample_sites": sample}


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C restatement; Arb cannot be built here) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pb = build_problem(args, 0, 0, use_engine=False)
    from oracle import c_port
    threads = host_threads()
    # size each step so that the whole run ends within a few minutes
    first = cpu_baseline(pb, target_seconds=2.0)
    rate = first["value"] / (pb["E"] * pb["C"])               # sites per second
    budget = 120.0 / max(1, args.steps + args.warmup)
    sample = int(min(pb["S"], max(1000, rate * min(budget, 15.0))))
    for _ in range(args.warmup):
        cpu_baseline(pb, sample=sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_baseline(pb, sample=sample)
    dt = time.perf_counter() - t0
    value = sample * pb["E"] * pb["C"] * args.steps / dt
    cb = {"value": value, "unit": "updates/s", "cores": threads, "kind": "port",
          "sample": "%d of %d site patterns per step, ll+deriv, oracle/c/plf_oracle.c (OpenMP, fp64)" % (sample, pb["S"])}
    line = {
        "impl": "reference", "metric": "site-edge-category updates/s (ll+deriv)", "value": value, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: GTR+Gamma4 (examples/BEAST.GTRG model) x %d-taxon Yule tree x %d site patterns "
                               "(bounded sample of %d per step), arbplf-ll + arbplf-deriv, site axis summed"
                               % (args.taxa, pb["S"], sample),
                   "note": "the Arb reference cannot be compiled in this image (no arb/flint/gmp/jansson headers); "
                           "this is its algorithm restated in C (oracle/c), fp64, one OpenMP thread per host core"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """Print the one JSON line on the real stdout (libraries such as NCCL write banners to fd 1)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # anything else that writes to fd 1 goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", type=int, default=1000000)
    ap.add_argument("--taxa", type=int, default=64)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
