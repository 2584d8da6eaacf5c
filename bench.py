#!/usr/bin/env python
"""bench.py -- headline benchmark: batched SOCPs solved per second (BASELINE.json).

A "step" is one pass of the hot path over one batch: the whole batched solve
(initial point + Mehrotra loop) of the workload's problems.  Workload at N=1 is
BASELINE.json configs[1] (C2): 10k random-feasible portfolio SOCPs, n=50,
1 LP block + 1 SOC of dim 51.  With N ranks every rank solves its own 10k
(weak scaling, no collective on the data path).

  value : problems/s with inputs already resident in HBM (socp_b200_solve_dev)
  e2e   : the same through the host-facing call: pinned host buffers -> set_data
          (H2D) -> solve -> get_results (D2H) inside the timed region
  roofline      : dominant kernel vs the FP64 peak measured in this run
  cpu_baseline  : the C oracle (port of the reference's dense path) on the host cores

`--impl reference` times the reference's CPU algorithm (the C oracle port; Julia
is not installed, so the reference itself cannot run) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p_ in (ROOT, os.path.join(ROOT, "socp.jl_b200")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

import numpy as np  # noqa: E402

METRIC = "batched SOCPs solved/sec"
UNIT = "problems/s"


def workload_meta(name):
    from socp_b200 import generators as gen
    cfg = gen.CONFIGS[name]
    if cfg["kind"] == "portfolio":
        n, p, k = cfg["n"], 1, 2 * cfg["n"] + 1
        desc = f"{name}: portfolio SOCPs n={n}, p=1, k={k} (POC {n} + SOC {n + 1}), batch {cfg['batch']}/GPU"
    else:
        n, p, k = cfg["n"], cfg["p"], cfg["ncones"] * cfg["dim"]
        desc = (f"{name}: random-feasible SOCPs n={n}, p={p}, k={k} ({cfg['ncones']} x SOC{cfg['dim']}), "
                f"batch {cfg['batch']}/GPU")
    return n, p, k, cfg["batch"], desc


def config_dict(name, problems_per_gpu):
    """The `config` object of the JSON line: identical in both arms (`--impl b200` and `--impl reference`) for the same
    command line; everything measured goes under `detail`."""
    n, p, k, batch, desc = workload_meta(name)
    bytes_step = problems_per_gpu * (8.0 * (k * n + p * n + n + p + k) + 8.0 * (n + p + 2 * k) + 24.0)
    return {"workload": desc, "problems_per_step_per_gpu": int(problems_per_gpu),
            "l2": (f"inputs {bytes_step / 1e6:.0f} MB per step exceed the 126 MB L2 (no flush needed)" if bytes_step > 130e6
                   else "inputs fit in L2; the device-resident leg re-reads them from L2/HBM every step (no flush)")}


def flops_per_iteration(n, p, k):
    """factor n(n+1)k + n^3/3 (+p terms), solve 2n^2 + 4nk, residuals 4nk + 4pn."""
    ff = n * (n + 1) * k + n ** 3 / 3.0 + (p * n * n + p * p * n + p ** 3 / 3.0 if p else 0.0)
    fs = 2 * n * n + 4 * n * k + (4 * p * n + 2 * p * p if p else 0.0)
    fr = 4 * n * k + 4 * p * n
    return ff + 2 * fs + fr


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.15)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx.append(float(s[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_fp64_peak(torch, dev, sustained_s=4.0):
    """FP64 GEMM peak of this GPU, measured in-run (cuBLAS DGEMM through torch.matmul; MEASURED_PEAKS.json has no FP64
    figure): burst = best of 10 at 8192^3, sustained = back to back for `sustained_s` seconds."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    burst = 2.0 * n ** 3 / (best * 1e-3) / 1e12
    reps = max(4, int(sustained_s / (best * 1e-3)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize(dev)
    sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    return burst, sustained


def cpu_baseline_run(name, nproblems, nthreads):
    """The C oracle (port of the reference's dense path) over `nproblems` of the workload."""
    from socp_b200 import generators as gen
    from oracle import c_oracle as co
    prob = gen.make_config(name, batch=nproblems)
    cones = tuple((c.kind, c.offs, c.dim) for c in prob.cones)
    t0 = time.perf_counter()
    r = co.solve_batch(prob.c, prob.A_cm, prob.b, prob.G_cm, prob.h, cones, sing=prob.sing, nthreads=nthreads,
                       shared_A=prob.shared_A, shared_G=prob.shared_G)
    dt = time.perf_counter() - t0
    return nproblems / dt, dt, r


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm (C oracle port, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, p, k, batch, desc = workload_meta(args.config)
    from oracle import c_oracle as co
    threads = os.cpu_count() or 1
    sample = args.ref_sample
    cpu_baseline_run(args.config, min(sample, 64), threads)      # warm (build + page in)
    for _ in range(max(0, args.warmup - 1)):
        cpu_baseline_run(args.config, min(sample, 64), threads)
    times = []
    conv = 0
    for _ in range(args.steps):
        v, dt, r = cpu_baseline_run(args.config, sample, threads)
        times.append(dt)
        conv = int((r["status"] == 0).sum())
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.config, args.batch or batch),
        "detail": {"sample": f"{sample} problems per step (bounded sample of the workload)", "converged": conv},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} problems/step x {args.steps} steps, C oracle (oracle/socp_oracle.c) "
                                   f"= C restatement of Socp.jl's dense path; Julia is not installed so the reference "
                                   f"itself cannot run; converged {conv}/{sample}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU (default: the config's)")
    ap.add_argument("--path", default="auto", choices=["auto", "tiled", "fused"])
    ap.add_argument("--ref-sample", type=int, default=None, help="problems per reference-arm step (default: per config)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="problems of the cpu_baseline leg (default: per config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-input", default="auto", choices=["auto", "dense", "csc"],
                    help="host storage of A and G handed to the e2e call: the reference's SparseMatrixCSC when G has "
                         "structural zeros (auto), or always dense / always CSC")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    # bounded CPU samples of the workload: tens of core-seconds per pass of the C oracle (C2: the whole 10k batch)
    cpu_default = {"C2": 10000, "C3": 50000, "C4": 16, "C5": 1}.get(args.config, 64)
    args.ref_sample = args.ref_sample or cpu_default
    args.cpu_sample = args.cpu_sample or cpu_default

    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import socp_b200 as sb
    from socp_b200 import generators as gen
    from socp_b200 import _lib as L
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, p, k, batch_cfg, desc = workload_meta(args.config)
    B = args.batch or batch_cfg
    from socp_b200 import sharding
    numa_node = -1
    if world > 1 and not os.environ.get("SOCP_B200_NO_NUMA_BIND"):
        # one process per GPU: its host thread and the pinned staging buffers allocated below go to the GPU's NUMA node
        # (N = 1 is left alone: the cpu_baseline leg of that run needs every host core)
        pr_ = torch.cuda.get_device_properties(local)
        numa_node = sharding.bind_host_to_pci_device(f"{pr_.pci_domain_id:04x}:{pr_.pci_bus_id:02x}:{pr_.pci_device_id:02x}.0")
    # every rank generates its own shard: problems [rank*B, (rank+1)*B) of the seeded sequence
    first, _last = sharding.weak_shard(B, rank)
    prob = gen.make_config(args.config, batch=B, first=first)

    # pinned host staging for the e2e leg (inputs) and outputs
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    hc, hh, hG = pin(prob.c), pin(prob.h), pin(prob.G_cm)
    hA, hb = (pin(prob.A_cm), pin(prob.b)) if p else (None, None)
    hsing = pin(prob.sing) if prob.sing is not None else None
    ox = torch.empty((B, n), dtype=torch.float64).pin_memory()
    ostatus = torch.empty(B, dtype=torch.int32).pin_memory()
    oiters = torch.empty(B, dtype=torch.int32).pin_memory()
    opobj = torch.empty(B, dtype=torch.float64).pin_memory()
    odobj = torch.empty(B, dtype=torch.float64).pin_memory()

    def dptr(t):
        return C.cast(t.data_ptr(), L.c_double_p) if t is not None else None

    def iptr(t):
        return C.cast(t.data_ptr(), L.c_int32_p)

    ss = sb.SolverState(prob, devices=[local])
    h = ss.handle
    lib = h.lib
    prm = sb.default_params(path={"auto": 0, "tiled": 1, "fused": 2}[args.path])
    flags = (1 if prob.shared_A else 0) | (2 if prob.shared_G else 0)
    sing_ptr = C.cast(hsing.data_ptr(), L.c_uint8_p) if hsing is not None else None

    def upload():
        h.check(lib.socp_b200_set_data(h.ptr, dptr(hc), dptr(hA), dptr(hb), dptr(hG), dptr(hh), sing_ptr, flags), "set_data")

    def solve_dev():
        h.check(lib.socp_b200_solve_dev(h.ptr, C.byref(prm)), "solve_dev")

    def download():
        h.check(lib.socp_b200_get_results(h.ptr, dptr(ox), None, None, None, iptr(ostatus), iptr(oiters), dptr(opobj),
                                          dptr(odobj)), "get_results")

    def solve_host():
        # the reference-facing call: Problem(c,A,b,G,h,cones) + solve_socp(prob, ss) from HOST data to HOST results
        h.check(lib.socp_b200_solve_host(h.ptr, C.byref(prm), dptr(hc), dptr(hA), dptr(hb), dptr(hG), dptr(hh), sing_ptr,
                                         flags, dptr(ox), None, None, None, iptr(ostatus), iptr(oiters), dptr(opobj),
                                         dptr(odobj)), "solve_host")

    # The reference keeps A and G as SparseMatrixCSC (src/Socp.jl:25,29): when G has structural zeros the e2e leg hands
    # the library exactly that (pattern + stored values per problem, built here once, outside the timed region) through
    # socp_b200_solve_host_csc, so only the stored values cross PCIe.  --e2e-input dense forces the dense call.
    csc = None
    if args.e2e_input != "dense" and not prob.shared_G:
        patG = (prob.G_cm != 0).any(axis=0)                   # (n, k): column-major union pattern of the batch
        nnzG = int(patG.sum())
        if args.e2e_input == "csc" or nnzG < 0.75 * k * n:
            def to_csc(M_cm, pat):
                cols, rows = np.nonzero(pat)                  # column by column, rows ascending: CSC order
                colptr = np.concatenate([[0], np.cumsum(pat.sum(axis=1))]).astype(np.int64)
                vals = np.ascontiguousarray(M_cm[:, cols, rows])
                return colptr, rows.astype(np.int64), pin(vals)
            cpG, rvG, hvG = to_csc(prob.G_cm, patG)
            cscG = L.Csc(nnzG, cpG.ctypes.data_as(L.c_int64_p), rvG.ctypes.data_as(L.c_int64_p), dptr(hvG), 0)
            cscA, hvA = None, None
            if p:
                patA = (prob.A_cm != 0).any(axis=0)
                cpA, rvA, hvA = to_csc(prob.A_cm, patA)
                cscA = L.Csc(int(patA.sum()), cpA.ctypes.data_as(L.c_int64_p), rvA.ctypes.data_as(L.c_int64_p), dptr(hvA), 0)
            csc = dict(G=cscG, A=cscA, keep=(cpG, rvG, hvG, hvA, (cpA, rvA) if p else None))

    def solve_host_csc():
        h.check(lib.socp_b200_solve_host_csc(h.ptr, C.byref(prm), dptr(hc), C.byref(csc["A"]) if p else None, dptr(hb),
                                             C.byref(csc["G"]), dptr(hh), sing_ptr, flags, dptr(ox), None, None, None,
                                             iptr(ostatus), iptr(oiters), dptr(opobj), dptr(odobj)), "solve_host_csc")

    e2e_call = solve_host_csc if csc else solve_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if csc:
        h2d_bytes = sum(t.numel() * t.element_size() for t in (hc, hh, hb, hsing, csc["keep"][2], csc["keep"][3]) if t is not None)
    else:
        h2d_bytes = sum(t.numel() * t.element_size() for t in (hc, hh, hG, hA, hb, hsing) if t is not None)
    d2h_bytes = sum(t.numel() * t.element_size() for t in (ox, ostatus, oiters, opobj, odobj))

    fp64_peak, fp64_sustained = measure_fp64_peak(torch, dev)
    hbm_peak, hbm_src = measured_peaks()

    # ---- device-resident leg (`value`)
    upload()
    for _ in range(args.warmup):
        solve_dev()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        solve_dev()
        tm = h.timings()
        dev_ms += tm["solve_ms"]
        launches += tm["kernel_launches"]
    barrier()
    wall_dev = time.perf_counter() - t0
    download()
    status = ostatus.numpy().copy()
    iters = oiters.numpy().copy()
    path_used = h.timings()["path_used"]

    # ---- end-to-end leg (`e2e`): H2D of the step's inputs + solve + D2H of its results, every step
    for _ in range(2):
        e2e_call()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    barrier()
    wall_e2e = time.perf_counter() - t1
    status_e2e = ostatus.numpy().copy()
    assert np.array_equal(status_e2e, status), "e2e leg and device-resident leg disagree on status"
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # max over ranks of the timed regions
    tt = torch.tensor([dev_ms * 1e-3, wall_dev, wall_e2e], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(B), float((status == 0).sum()), float(iters.sum()), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_dev, t_wall_dev, t_e2e = [float(v) for v in tt.tolist()]
    total_B, total_conv, total_iters, total_launches = [float(v) for v in cnt.tolist()]

    if rank == 0:
        value = total_B * args.steps / t_dev
        e2e_value = total_B * args.steps / t_e2e
        # roofline of the dominant kernel (the whole-solve kernel on the fused path; the sum of the solve's
        # kernels on the tiled path): algorithmic FP64 flops = F_it x iterations actually taken
        f_it = flops_per_iteration(n, p, k)
        # +1: the initial point costs one factor + one solve (counted as one iteration's factor+solve share)
        flops_step = f_it * (total_iters + total_B) / world          # per GPU
        achieved = flops_step * args.steps / t_dev / 1e12             # t_dev: max over ranks
        bytes_step = B * (8.0 * (k * n + p * n + n + p + k) + 8.0 * (n + p + 2 * k) + 24.0)
        hbm_ach = bytes_step * args.steps / t_dev / 1e9
        # DRAM traffic of the dominant kernel from the latest `ncu --set full` capture (tools/gpu_profiles.sh writes
        # profiles/ncu_traffic.json with the kernel name and the git SHA it was taken at); null when it is not for
        # this workload / batch / path
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath)).get(args.config)
            if t and t.get("problems_per_launch") == B and path_used == 2:
                traffic = t["dram_bytes"]
                traffic_src = f"{t.get('kernel', '?')} @ {t.get('git_sha', '?')}"
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_baseline_run(args.config, min(args.cpu_sample, 64), threads)
            v, dt, r = cpu_baseline_run(args.config, args.cpu_sample, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample} problems of the same workload in {dt:.2f} s, C oracle "
                             f"(oracle/socp_oracle.c, OpenMP, one problem per thread)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.config, B),
            "detail": {"path": {1: "tiled", 2: "fused"}.get(path_used, "?"), "converged": int(total_conv),
                       "mean_iters": total_iters / total_B,
                       # BASELINE.json's secondary metric: one "KKT factor+solve" = one setup_iter + one solve_kkt; a
                       # solve does (iterations + 1) of them per problem (the initial point is one).  Amortised over the
                       # whole job (throughput), not the latency of one problem.
                       "us_per_kkt_factor_solve_amortised": 1e6 * t_dev / args.steps / (total_iters + total_B),
                       "timing": "CUDA events on the library's launch stream, summed over the timed steps, max over ranks",
                       "wall_s_device_leg": t_wall_dev},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "input": ("A and G as SparseMatrixCSC (the reference's storage, src/Socp.jl:25,29): %d of %d entries of G "
                              "stored" % (csc["G"].nnz, k * n)) if csc else "A and G dense column-major",
                    "how": ("socp_b200_solve_host_csc" if csc else "socp_b200_solve_host") +
                           ": pinned host buffers -> chunked H2D / (CSC scatter) / solve / D2H (overlapped on three "
                           "streams) -> pinned host results; wall clock between barriers, max over ranks",
                    "host_numa_node_rank0": numa_node},
            "gpu_launches": int(total_launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "FP64 burst: cuBLAS DGEMM 8192^3 through torch.matmul, best of 10, measured in this "
                                        "run (MEASURED_PEAKS.json has no FP64 figure); the kernel is timed alone, so burst",
                         "peak_sustained": fp64_sustained, "frac_of_sustained": achieved / fp64_sustained,
                         "flops_note": "algorithmic flops of the dense formulation (SURVEY.md 8d) x iterations taken; "
                                       "the kernel skips the structural zeros of G (singleton / empty rows)",
                         "flops_per_step": flops_step, "hbm_achieved_gbs": hbm_ach, "hbm_peak_gbs": hbm_peak,
                         "hbm_frac": hbm_ach / hbm_peak, "hbm_peak_source": hbm_src},
            "clocks": sampler.summary(),
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
