#!/usr/bin/env python
"""bench.py — TT-cross sweep throughput on B200 (contract in the task description).

A "step" is one complete greedy TT-cross integration (one `dtt_dmrgg` call, reference lib/dmrgg.f90:11) of the
workload BASELINE.json's metric is quoted on: `test_crs_ising c 10 256 32 2` (Ising C_10, n = 257, maxrank 32,
rook depth 2).  The bonds are cut into P = 8 partitions (the reference's MPI partition; results are a function of
the partition, not of the GPU count) which are mapped block-wise onto the N GPUs.

metric : integrand evaluations per second = neval / device time of the step (max over ranks)
e2e    : the same through the C-ABI with HOST buffers: par + quad upload + sweep + copy-out of all cores + integral every step
         (handle created once; `cold_ms_per_step` adds ttc_create / ttc_destroy)
--impl reference : the CPU oracle (restatement of the reference; the Fortran reference cannot be built here) on all host
                   threads, same workload, same partition.
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
sys.path.insert(0, ROOT)

# BASELINE.json configs: name -> (family, driver arguments, RANK, PIV, partitions, label).  B is the headline (configs[1]).
CONFIGS = {
    "A": ("ising", ("c", 6, 64), 16, 1, 4, "test_crs_ising c 6 64 16 1"),
    "B": ("ising", ("c", 10, 256), 32, 2, 8, "test_crs_ising c 10 256 32 2"),
    "C": ("ising", ("d", 8, 256), 48, 2, 6, "test_crs_ising d 8 256 48 2"),
    "D": ("ising", ("e", 6, 512), 64, 3, 4, "test_crs_ising e 6 512 64 3"),
    "E": ("mvn", (64, 128), 32, 1, 63, "test_crs_mvn 64 128 32 (pivoting 1)"),
}


L2_NOTE = "GPU arm: L2 flushed between timed steps (256 MiB write); working set < 126 MB L2"


def make_problem(T, name):
    fam, a, R, piv, P, label = CONFIGS[name]
    prob = T.drivers.ising(*a) if fam == "ising" else T.drivers.mvn(*a)
    return prob, R, piv, P, label


def flops_per_eval(fam, a, d):
    """operations of one integrand evaluation (Ising C 5d+3; D/E add the d(d+1)/2 pair products; MVN 3d^2 + d)."""
    if fam == "mvn":
        return 3 * d * d + d
    return 5 * d + 3 if a[0] == "c" else 5 * d + 3 + 5 * (d * (d + 1) // 2)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons of ONE GPU, polled every 50 ms from before the warm-up until after the timed
    regions; stop() keeps the samples whose host time stamp falls inside the timed regions."""

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=index,clocks.sm,clocks.max.sm,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [(ts, r) for ts, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for ts, r in rows if t_begin is not None and t_begin - 0.05 <= ts <= t_end + 0.05]
        window = "timed regions"
        if not inside:                       # (a timed region shorter than the polling period: take the neighbours)
            inside = [r for ts, r in rows]
            window = "whole run (no sample fell inside the timed regions)"
        sm = sorted(float(r[1]) for r in inside)
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(inside[0][2]), "reasons": sorted(reasons), "samples": len(inside),
                "window": window}


def reference_arm(args):
    """CPU arm: the oracle (a restatement of the reference, kind = "port": the Fortran reference cannot be built -- there is no
    gfortran / mpif90 / BLAS in the image or on the GPU box, profiles/r02_toolchain_probe.txt) with every host core: the
    virtual ranks run concurrently like the reference's MPI ranks, each with an OpenMP team for its evaluation loops."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    fam, a, R, piv, P, label = CONFIGS[args.config]
    s = O.ising_setup(*a) if fam == "ising" else O.mvn_setup(*a)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.lib().tto_set_num_threads(ncpu)          # explicit: torchrun exports OMP_NUM_THREADS=1 to its children
    O.lib().tto_set_rank_concurrency(1)
    cores = O.lib().tto_num_threads()
    o = O.Oracle(s)
    for _ in range(args.warmup):
        o.run(maxrank=R, piv=piv, P=P)
    t0 = time.perf_counter()
    neval = 0
    for _ in range(args.steps):
        r = o.run(maxrank=R, piv=piv, P=P)
        neval += r.neval
    dt = time.perf_counter() - t0
    v = neval / dt
    line = {
        "impl": "reference", "metric": "integrand_evals_per_s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "partitions": P, "neval_per_step": neval // args.steps, "l2": L2_NOTE},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full runs of the workload (CPU restatement of dtt_dmrgg: {P} virtual ranks side by side, "
                                   f"OpenMP inside each, {cores} threads in all; the Fortran reference cannot be built: no gfortran/MPI/BLAS)"},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-superblock", action="store_true")
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS), help="BASELINE.json configuration (B = the headline workload)")
    args = ap.parse_args()
    if args.impl == "reference":
        # the CPU arm at its best: idle OpenMP workers spin between the (many, short) parallel regions instead of sleeping --
        # +12 % on the GPU box's 16 cores (tests/_ref_arm_probe.py: 1.70e7 -> 1.92e7 evals/s; pinning threads halves it).  Must be
        # in the environment before libgomp initialises, i.e. before the oracle library is loaded.
        os.environ.setdefault("OMP_WAIT_POLICY", "active")
        return reference_arm(args)

    import numpy as np
    import ttcross_b200 as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version banner there) go to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl")
        dist = dist_mod

    fam, cargs, _, _, _, _ = CONFIGS[args.config]
    prob, R, piv, PARTITIONS, label = make_problem(T, args.config)
    PARTITIONS = max(PARTITIONS, world)
    hbm_peak, peak_src = _peaks()

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-input arm: handle and device buffers exist, a step is one ttc_dmrgg call
    t = prob.make(device=local_rank)
    t.set_partition(PARTITIONS)
    if world > 1:
        # the P = 8 core blocks are mapped block-wise onto the N ranks; the library's own NCCL communicator carries the
        # per-sweep exchange (pivot tape + boundary fibers + quadrature chains); torch.distributed only hands over its id
        T.multi.attach(t, dist)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                    # polling runs from before the warm-up: nvidia-smi needs a moment to start
    for _ in range(max(args.warmup, 3)):
        g = t.dmrgg(R, prob.accuracy, piv)
    if rank == 0:
        sampler.wait_first()
    launches0 = t.launch_count()
    barrier()
    t_timed_begin = time.perf_counter()
    dev_ms, sweep_ms = [], []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        t.l2_flush()                       # cold L2 between timed iterations (working set < 126 MB L2)
        g = t.dmrgg(R, prob.accuracy, piv)
        dev_ms.append(g.device_ms)
        sweep_ms.append(t.sweep_kernel()[1])           # CUDA events on the library's stream around the persistent kernel
    persistent, _, sweep_cluster, sweep_threads = t.sweep_kernel()
    barrier()
    wall = time.perf_counter() - wall0
    launches = t.launch_count() - launches0
    ms_step = float(np.mean(dev_ms))
    if dist is not None:
        import torch
        x = torch.tensor([ms_step], device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        ms_step = float(x.item())
    value = g.neval / (ms_step * 1e-3)

    # ---- e2e arm: host buffers in, host buffers out, every step.  The handle (device plan, captured graphs, pinned staging)
    # is created once, like the reference's mpi_init + allocation of `tt`; every step ships par + quad from host memory
    # (ttc_set_par / ttc_set_quad -> uploaded by the next ttc_dmrgg), runs the cross, copies every core back and reads the
    # integral.  `cold_ms_per_step` is the same with ttc_create / ttc_destroy inside the timed region as well.
    e2e_t, cold_t = [], []
    # caller-owned result buffer, like arg%u(k)%p -- sized for maxrank and BOUND to the handle (ttc_bind_cores): the reference
    # returns the cores inside the dtt_dmrgg call (`arg` is inout, lib/dmrgg.f90:11-26), so every ttc_dmrgg below ends with
    # the device -> host transfer of the cores into this buffer
    host_out = np.empty(t.cores_capacity(R))
    t.bind_cores(host_out)
    h2d = prob.par.nbytes + prob.quad.nbytes + prob.n.nbytes
    d2h = 0
    for it in range(args.steps + 1):
        barrier()
        t0 = time.perf_counter()
        t.set_par(prob.par)
        t.set_quad(prob.quad)
        ge = t.dmrgg(R, prob.accuracy, piv)
        cores = t.cores(out=host_out)                    # device -> host: every core this rank holds
        val = t.quad()
        barrier()
        dt = time.perf_counter() - t0
        d2h = sum(c.nbytes for c in cores) + 8 + ge.pivlog.nbytes
        if it > 0:
            e2e_t.append(dt)
    t.bind_cores(None)
    if world == 1:
        for it in range(3):
            t0 = time.perf_counter()
            te = prob.make(device=local_rank)            # ttc_create + first upload
            te.set_partition(PARTITIONS)
            te.dmrgg(R, prob.accuracy, piv)
            te.cores(); te.quad()
            te.close()
            if it > 0:
                cold_t.append(time.perf_counter() - t0)
    e2e_mean = float(np.mean(e2e_t))
    if dist is not None:
        import torch
        x = torch.tensor([e2e_mean, float(d2h)], device="cuda", dtype=torch.float64)
        y = x.clone()
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        e2e_mean = float(x[0].item())
        d2h = int(y[1].item())                           # all ranks' copies
        h2d *= world
    e2e_val = g.neval / e2e_mean
    clocks = sampler.stop(t_timed_begin, time.perf_counter()) if rank == 0 else None

    # ---- roofline of the dominant kernel: per-class device times from a profiled pass (events around every launch)
    t.set_profile(True)                                  # collective like every ttc_dmrgg call when world > 1
    gp = t.dmrgg(R, prob.accuracy, piv)
    prof = t.profile()
    t.set_profile(False)
    tot_ms = sum(v[1] for v in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    # dominant kernel of the step: k_visits (ttc_visit.cuh) — one launch per sweep runs every bond visit of this GPU's
    # partitions.  ALGORITHMIC bytes per launch (DESIGN.md "k_visits"): every lottery candidate reads 2*r(p) factor values,
    # every fiber element reads r(p) factor values and writes fiber + residual (16 B), the rank-1 append re-reads and
    # writes the last column and row fibers (32 B per element).  Ranks at sweep `it` are min(it, final rank) (one pivot per
    # bond and sweep); all 2*piv rook steps are counted (upper bound), averaged over the sweeps of the step.
    ranks = g.ranks
    nn = int(prob.n[0])
    own = T.multi.share(1, prob.d - 1, PARTITIONS)
    v0, v1 = T.multi.block_of(PARTITIONS, world, rank)
    my_bonds = list(range(int(own[v0]), int(own[v1])))
    fpe = flops_per_eval(fam, cargs, prob.d)
    byt, flo, evs = [], [], []
    for it in range(1, int(g.nsweeps) + 1):
        b = f = 0.0
        e_ = 0
        for pb in my_bonds:
            r0, r1, r2 = (min(it, int(ranks[q])) for q in (pb - 1, pb, pb + 1))
            nlot, ncol, nrow = r0 + 2 * nn + r2, r0 * nn, nn * r2
            ev = nlot + max(piv, 1) * (ncol + nrow)
            b += nlot * 16.0 * r1 + max(piv, 1) * (ncol + nrow) * (8.0 * r1 + 16.0) + (ncol + nrow) * 32.0
            f += ev * (fpe + 2 * r1)
            e_ += ev
        byt.append(b)
        flo.append(f)
        evs.append(e_)
    ncu = {}
    for fn in ("r02_ncu_summary.json", "r01_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as fjs:
                ncu = json.load(fjs)
            break
        except Exception:
            pass
    pk_nofma_bench = None
    try:
        pk_nofma_bench = T.fp64_peak(local_rank, False)
    except Exception:
        pass
    if persistent:
        # dominant kernel = the persistent sweep kernel k_sweeps (ttc_sweep.cuh): ONE launch runs every sweep of the step
        dom_ms = float(np.mean(sweep_ms))
        per_launch_bytes, per_launch_flops = float(np.sum(byt)), float(np.sum(flo))
        strict_bytes = 8.0 * float(np.sum(evs))
        achieved = per_launch_bytes / (dom_ms * 1e-3) / 1e9
        tflops = per_launch_flops / (dom_ms * 1e-3) / 1e12
        roofline = {"kernel": f"k_sweeps (persistent cooperative kernel, {sweep_cluster} CTAs x {sweep_threads} threads per partition: all sweeps of the step -- "
                              "TMA-staged pivot tables, lottery, rook fibers + residuals, cluster argmax folds, rank-1 append, neighbour exchange, exit test)",
                    "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": ncu.get("k_sweeps", {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": per_launch_bytes,
                    "algorithmic_bytes_per_launch_strict_8B_per_evaluation": strict_bytes, "frac_strict": strict_bytes / (dom_ms * 1e-3) / 1e9 / hbm_peak,
                    "algorithmic_flops_per_launch": per_launch_flops, "achieved_tflops": tflops,
                    "fp64_frac_of_dmul_dadd_peak": (tflops / pk_nofma_bench) if pk_nofma_bench else None,
                    "share_of_step": dom_ms / ms_step if ms_step else None, "avg_launch_us": 1e3 * dom_ms, "launches_per_step": 1,
                    "timing": "CUDA events recorded by the library on its own stream around the launch (ttc_sweep_kernel_ms), mean over the timed steps",
                    "note": "latency-bound by construction: a sweep is a chain of ~6 dependent steps of <= r*n evaluations per partition, each ending in a "
                            "cluster-wide first-index argmax (SURVEY F4); the factors stay in the 126 MB L2, so neither roofline binds -- the algorithmic "
                            "bytes count the factor values every residual reads (8 r(p) + 16 B per fiber element), the strict figure 8 B per kept evaluation; "
                            "the roofline-sized kernel of the path is reported in roofline_superblock"}
    else:
        dom_name = "bond_visits_cluster" if prof.get("bond_visits_cluster", (0, 0))[0] else "fiber_eval_residual"
        fl, fms = prof[dom_name]
        dom_avg_ms = fms / max(fl, 1)
        per_launch_bytes = float(np.mean(byt)) * (1.0 if dom_name == "bond_visits_cluster" else 1.0 / (2 * max(piv, 1) + 1))
        achieved = per_launch_bytes / (dom_avg_ms * 1e-3) / 1e9
        roofline = {"kernel": "k_visits (cluster per partition, one launch per sweep)" if dom_name == "bond_visits_cluster" else "k_fiber",
                    "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": ncu.get("k_visits", {}).get("dram_bytes_per_launch") if dom_name == "bond_visits_cluster" else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch_bytes, "algorithmic_flops_per_launch": float(np.mean(flo)),
                    "share_of_step": fms / tot_ms if tot_ms else None, "avg_launch_us": 1e3 * dom_avg_ms,
                    "note": "per-sweep schedule (the persistent kernel did not fit this configuration): latency-bound, factors L2-resident"}

    line = {
        "metric": "integrand_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        # (the same dict in both arms, so the driver's config comparison holds; what is specific to this run is in `details`)
        "config": {"workload": label, "partitions": PARTITIONS, "neval_per_step": int(g.neval), "l2": L2_NOTE},
        "details": {"partition_map": (f"{PARTITIONS} core blocks block-mapped onto {world} GPU(s); per-sweep exchange = stores into the neighbours' peer-memory "
                                      "windows (CUDA IPC over NVLink) + release/acquire flags inside the persistent kernel; NCCL only hands out the window handles")
                    if world > 1 else f"{PARTITIONS} core blocks, one cluster each, in one persistent kernel on one GPU",
                    "schedule": "persistent kernel" if persistent else "per-sweep kernels (CUDA graph)",
                    "sweeps": int(g.nsweeps), "final_ranks": [int(x) for x in g.ranks], "integral": float(g.vals[-1])},
        "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_mean, "cold_ms_per_step": (1e3 * float(np.mean(cold_t)) if cold_t else None),
                "handle": "reused across steps (ttc_set_par + ttc_set_quad ship the inputs every step; the cores land in a caller-owned host buffer bound with ttc_bind_cores, transferred inside every ttc_dmrgg)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "per_sweep_schedule_kernel_classes_ms": {k: {"launches": v[0], "ms": v[1]} for k, v in prof.items() if v[0]},
        "per_sweep_schedule_note": "diagnostic pass with an event pair around every launch; it runs the per-sweep schedule (k_visits + k_exchange_fused + ...), not the persistent kernel",
        "wall_s_timed_region": wall,
    }

    # ---- superblock kernel (pivoting = -1 branch, dmrgg.f90:341-396) at the workload's shape: the roofline-sized kernel
    if not args.no_superblock and rank == 0 and world == 1:
        try:
            bond = prob.d // 2
            r1 = int(ranks[bond])
            pk_fma, pk_nofma = T.fp64_peak(local_rank, True), T.fp64_peak(local_rank, False)
            sb = t.superblock_probe(bond, store=False, reps=5, variant=0)
            sbf = t.superblock_probe(bond, store=False, reps=5, variant=2)
            sbs = t.superblock_probe(bond, store=True, reps=5, variant=0)
            sbp = t.superblock_probe(bond, store=False, reps=2, variant=1)
            try:
                sbm = t.superblock_probe(bond, store=False, reps=5, variant=3)
            except Exception:
                sbm = None
            flops = sb["count"] * (fpe + 2 * r1)          # operations of an evaluation + 2 r(p) per residual element
            def tf(ms):
                return flops / (ms * 1e-3) / 1e12
            line["roofline_superblock"] = {
                "kernel": "k_superblock_t (ttc_superblock.cuh): evaluate a(i,j,k,q), residual against col*row (K = r(p)), two first-index argmaxes",
                "shape": [int(ranks[bond - 1]), nn, nn, int(ranks[bond + 1])], "K": r1, "elements": sb["count"],
                "algorithmic_flops_per_launch": flops,
                "fp64_peak_measured_tflops": {"dfma": pk_fma, "dmul_dadd": pk_nofma,
                                              "how": "ttc_fp64_peak: 8 independent chains per thread, 2 x 512 threads per SM, CUDA events"},
                "fused": {"ms": sb["ms"], "bound": "fp64", "achieved": tf(sb["ms"]), "unit": "TFLOP/s", "peak": pk_nofma,
                          "frac": tf(sb["ms"]) / pk_nofma, "frac_of_dfma_peak": tf(sb["ms"]) / pk_fma,
                          "evals_per_s": sb["count"] / (sb["ms"] * 1e-3),
                          "traffic": ncu.get("k_superblock_t", {}).get("dram_bytes_per_launch"),
                          "note": "reference arithmetic (no FMA contraction): the DMUL+DADD ceiling is the bound; bit-identical to the oracle"},
                "fused_dfma": {"ms": sbf["ms"], "achieved": tf(sbf["ms"]), "unit": "TFLOP/s", "peak": pk_fma, "frac": tf(sbf["ms"]) / pk_fma,
                               "note": "residual contracted into DFMA: not bit-exact, never used by the sweep"},
                "stored": {"ms": sbs["ms"], "bound": "hbm", "achieved": 8.0 * sbs["count"] / (sbs["ms"] * 1e-3) / 1e9, "unit": "GB/s",
                           "peak": hbm_peak, "frac": 8.0 * sbs["count"] / (sbs["ms"] * 1e-3) / 1e9 / hbm_peak,
                           "note": "also writes a (8 B per element); still FP64-bound at this shape"},
                "plain_kernel_ms": sbp["ms"],
                "fast_dmma": ({"ms": sbm["ms"], "achieved": tf(sbm["ms"]), "unit": "TFLOP/s", "peak": pk_fma, "frac": tf(sbm["ms"]) / pk_fma,
                               "same_residual_argmax_as_parity_mode": bool(sbm["argmax_b"] == sb["argmax_b"]),
                               "residual_value_rel_diff": abs(sbm["b"] / sb["b"] - 1.0) if sb["b"] else None,
                               "note": "fast mode: residual through mma.sync.m8n8k4.f64 (SASS DMMA), evaluated tile parked in shared memory; measured SLOWER than "
                                       "the vector path -- on B200 the FP64 tensor rate equals the DFMA rate (profiles/r02_ubench.txt: 37.06 vs 36.89 TFLOP/s) and "
                                       "the kernel is bound by the evaluation's instruction mix, not by the contraction"} if sbm else None),
                "other_shapes": "profiles/r02_superblock_shapes.txt (C, D, E: generic per-element evaluation, 10-12 % of the DMUL+DADD ceiling)",
            }
        except Exception as e:  # noqa: BLE001
            line["roofline_superblock"] = {"error": str(e)}

    # ---- the tall-skinny QR of ort0_d (lib/ort.f90:17-81; SURVEY 8 row a21, north-star item 3) at this config's unfolding shape
    # (r n) x r: not called by the sweep -- reported beside it, against the measured FP64 ceilings
    if not args.no_superblock and rank == 0 and world == 1:
        try:
            mq, nq = R * nn, R
            rngq = np.random.default_rng(7)
            aq = np.asfortranarray(rngq.standard_normal((mq, nq)) * np.exp(rngq.uniform(-3, 3, size=(1, nq))))
            qq_, rq_, ms_q = T.qr_thin(aq, device=local_rank, reps=10)
            fl_q = 4.0 * mq * nq * nq - 4.0 * nq ** 3 / 3.0            # dgeqrf + dorgqr
            pk_nofma_q = line.get("roofline_superblock", {}).get("fp64_peak_measured_tflops", {}).get("dmul_dadd") or T.fp64_peak(local_rank, False)
            line["roofline_qr"] = {
                "kernel": "TSQR of ttc_qr.cuh (k_tsqr_factor per tree level, k_tsqr_formq, k_tsqr_leaf, k_tsqr_sign, k_tsqr_scale): thin Householder QR with LAPACK's signs, explicit Q",
                "shape": [mq, nq], "ms": ms_q, "algorithmic_flops": fl_q, "achieved": fl_q / (ms_q * 1e-3) / 1e12, "unit": "TFLOP/s",
                "peak": pk_nofma_q, "frac": fl_q / (ms_q * 1e-3) / 1e12 / pk_nofma_q, "bound": "latency",
                "hbm_frac": 2.0 * 8.0 * mq * nq / (ms_q * 1e-3) / 1e9 / hbm_peak,
                "residual_vs_lapack": float(np.abs(rq_ - np.linalg.qr(aq)[1]).max() / np.linalg.norm(aq)),
                "note": "n dependent reflectors per tree level (sum of squares, sqrt, two divisions, rank-1 update each): latency-bound by construction; "
                        "DMMA does not apply -- the only dense contraction (Q1 * M per block) is < 10 % of the time (profiles/r02f_ncu_tsqr.txt)"}
        except Exception as e:  # noqa: BLE001
            line["roofline_qr"] = {"error": str(e)}

    # ---- CPU baseline beside it: the oracle on the host cores, bounded sample
    if not args.no_cpu_baseline and rank == 0 and args.gpus == 1:
        from oracle import oracle as O
        s = O.Setup(prob.kind, prob.d, prob.n, prob.par, prob.aux, prob.quad, prob.accuracy, prob.tru)
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        O.lib().tto_set_num_threads(ncpu)
        O.lib().tto_set_rank_concurrency(1)
        o = O.Oracle(s)
        t0 = time.perf_counter()
        reps = 0
        ne = 0
        while time.perf_counter() - t0 < 10.0 and reps < 20:
            r = o.run(maxrank=R, piv=piv, P=PARTITIONS)
            ne += r.neval
            reps += 1
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(r.pivlog, g.pivlog) and np.array_equal(r.vals, g.vals))
        line["cpu_baseline"] = {"value": ne / dt, "unit": "evals/s", "cores": O.lib().tto_num_threads(), "kind": "port",
                                "sample": f"{reps} full runs of the same workload and partition in {dt:.1f} s", "matches_gpu_bitwise": same}
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    t.close()            # collective when a communicator is attached (peer windows are unmapped before they are freed)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
