#!/usr/bin/env python
"""bench.py -- throughput of the nonbonded hot path on synthetic FCC Lennard-Jones fluids.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c5]

A "step" is one velocity-Verlet step of the whole fluid: [second half-kick + first half-kick + drift]
kernel, re-binning when an atom has moved more than skin/2 since the last binning (or at a fixed cadence,
--rebin-every), one LJ force evaluation from the pair list.  The timed region holds
exactly K steps issued as ONE emdee_vv_step call (no host synchronisation inside), bracketed by a
barrier and a device synchronisation, timed with CUDA events on the library's stream, max over ranks.

metric  : LJ pair-interactions/s = (unique pairs i<j with r2 <= rc2, counted by the audit kernel) x K / t
          ("atom_steps_per_s" is printed beside it: N x K / t) -- BASELINE.json's two headline numbers.
workload: c3 = LJ fluid N=4,000,000 (fcc 100^3), rc=2.5 sigma, rs=2.0, rho*=0.8442 -- the configuration
          BASELINE.json's target is quoted on; it fits one B200, and is strong-scaled over 1/2/4/8 GPUs.
e2e     : the same metric through the C ABI with HOST buffers: positions in pinned host memory ->
          emdee_set_positions (H2D) -> emdee_bin -> emdee_compute_nonbonded(F|E|V) -> forces, energies,
          virials back to pinned host memory (D2H), every iteration inside the timed region.
roofline: dominant kernel k_force_list_p (the pair-list stepping kernel, one launch per step); achieved =
          71 flop x pairs per launch / mean launch duration (CUDA events around every launch on the library's
          stream, emdee_profile_begin/end/kind); peak = DFMA throughput measured on this GPU in the same run
          (MEASURED_PEAKS.json holds no FP64 number).  The pair-list build kernel (one launch per re-binning)
          is reported beside it; "traffic" is the DRAM bytes per launch of the ncu capture under profiles/.
The oracle (oracle/) is used here only for the cpu_baseline leg and for --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_PAIR = 71
BYTES_FORCE_EVAL = 80          # B/atom: read x,y,z + LJ params, write f, e, w
BYTES_LIST_STEP = 280          # B/atom-step of the pair-list stepping kernel: 176 list + 54 recipe + 24 positions + 24 forces;
                               # the fused velocity-Verlet update adds BYTES_VV (+ r_bin, mass) to the same kernel
BYTES_VV = 120                 # B/atom-step: read r,v,f, write r,v
BYTES_REBIN = 72               # B/atom-rebin

WORKLOADS = {
    #        n (fcc cells/dim), cutoff, switch, description
    "c1": (10, 2.5, 2.0, "LJ fluid N=4000 (fcc 10^3), rc=2.5, rho*=0.8442"),
    "c2": (40, 2.5, 2.0, "LJ fluid N=256000 (fcc 40^3), rc=2.5, rho*=0.8442"),
    "c3": (100, 2.5, 2.0, "LJ fluid N=4000000 (fcc 100^3), rc=2.5, rho*=0.8442"),
    "c5": (200, 3.0, 2.5, "LJ fluid N=32000000 (fcc 200^3), rc=3.0, rho*=0.8442"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--ndiv", type=int, default=1)
    ap.add_argument("--skin", type=float, default=0.45)
    ap.add_argument("--rebin-every", type=int, default=-1,
                    help="steps between re-binnings; -1: adaptive (re-bin when an atom has moved more than skin/2)")
    ap.add_argument("--dt", type=float, default=0.005)
    ap.add_argument("--temperature", type=float, default=1.44)
    ap.add_argument("--e2e-iters", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args):
    import emdee_jl_b200 as em

    n, rc, rs, desc = WORKLOADS[args.workload]
    pos, L = em.workloads.fcc_lattice(n)
    N = pos.shape[0]
    return dict(n=n, rc=rc, rs=rs, desc=desc, pos=pos, L=L, N=N, atoms=em.workloads.lj_fluid_atoms(N),
                vel=em.workloads.maxwell_velocities(N, args.temperature))


def cpu_leg(args, w, budget_s, steps, warmup):
    """Times the CPU oracle (OpenMP, all host cores) on a bounded sample of the workload: velocity-Verlet
    steps of a same-density, same-cutoff FCC fluid small enough to finish in `budget_s`."""
    import emdee_jl_b200 as em
    from oracle import oracle_c

    oracle_c.build()
    cores = oracle_c.num_threads(fast=True)
    # probe the evaluation rate on config 2's size, then pick the largest sample that fits the budget
    pos, L = em.workloads.fcc_lattice(16)
    at = em.workloads.lj_fluid_atoms(pos.shape[0])
    oracle_c.cutoff_cells(pos, L, w["rc"], w["rs"], at, ndiv=args.ndiv, bitmask=1, fast=True)
    t0 = time.perf_counter()
    r = oracle_c.cutoff_cells(pos, L, w["rc"], w["rs"], at, ndiv=args.ndiv, bitmask=1, fast=True)
    rate = pos.shape[0] / (time.perf_counter() - t0)          # atom-evaluations/s, O(N)
    n = w["n"]
    while n > 8 and 4 * n ** 3 * (steps + warmup + 1) / rate > budget_s:
        n = max(8, int(n * 0.8))
    pos, L = em.workloads.fcc_lattice(n)
    N = pos.shape[0]
    at = em.workloads.lj_fluid_atoms(N)
    vel = em.workloads.maxwell_velocities(N, args.temperature)
    mass = np.ones(N)
    r = oracle_c.cutoff_cells(pos, L, w["rc"], w["rs"], at, ndiv=args.ndiv, bitmask=1, fast=True)
    p, v, f = pos, vel, r["forces"]
    if warmup:
        p, v, f = oracle_c.vv_steps(p, v, f, mass, L, w["rc"], w["rs"], at, args.dt, warmup, ndiv=args.ndiv, fast=True)
    t0 = time.perf_counter()
    p, v, f = oracle_c.vv_steps(p, v, f, mass, L, w["rc"], w["rs"], at, args.dt, steps, ndiv=args.ndiv, fast=True)
    dt = time.perf_counter() - t0
    npairs = oracle_c.cutoff_cells(p, L, w["rc"], w["rs"], at, ndiv=args.ndiv, bitmask=1, fast=True)["npairs"]
    return dict(value=npairs * steps / dt, unit="pair-interactions/s", cores=cores, kind="port",
                sample="%d velocity-Verlet steps of an N=%d fcc LJ fluid (same rho*, rc, rs, dt; full-neighbour "
                       "OpenMP cell-list oracle, -O3 AVX2+FMA)" % (steps, N),
                atom_steps_per_s=N * steps / dt, ms_per_step=dt / steps * 1e3, N=N)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia
    (not installed; nothing under /root/reference compiles with gcc), so oracle/_ref does not exist and
    this arm times the oracle port on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(zip(("n", "rc", "rs", "desc"), WORKLOADS[args.workload]))
    r = cpu_leg(args, w, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "LJ pair-interactions/s", "value": r["value"], "unit": "pair-interactions/s",
        "atom_steps_per_s": r["atom_steps_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": r["sample"], "ndiv": args.ndiv, "dt": args.dt},
        "cpu_baseline": {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    import emdee_jl_b200 as em

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = em.Context(local)
    if world > 1:
        ids = [em.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])

    if world > 1 and args.rebin_every < 0:
        # the adaptive criterion costs an all-reduce and a host read-back per step: at 8 GPUs (0.5 ms steps) that is more
        # than the re-binnings it saves (measured: 0.65 vs 0.49 ms/step), so slab runs re-bin at a fixed, safe cadence
        args.rebin_every = 5
    w = workload(args)
    N, L = w["N"], w["L"]
    s = em.NonbondedSystem(N, L, ctx)
    s.set_model(em.LennardJonesModel(w["rc"], w["rs"]))
    s.set_atoms(w["atoms"])
    s.set_positions(w["pos"])
    s.set_velocities(w["vel"])
    s.set_masses(np.ones(N))
    s.set_skin(args.skin)
    s.bin(args.ndiv)
    s.compute(em.CUTOFF, em.FORCES)

    def barrier():
        s.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fp64_peak = ctx.measure_fp64_peak()

    # ---- warm-up, then exactly K timed steps -----------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # samples cover warm-up + timed region (the timed region alone can be < 100 ms)
    s.vv_step(args.dt, args.warmup, args.rebin_every)
    pairs0 = allsum(float(s.pair_set_digest()[0]))
    s.compute(em.CUTOFF, em.FORCES)
    s.vv_step(args.dt, args.warmup, args.rebin_every)
    barrier()
    launches0 = ctx.launch_count()
    s.profile_begin()
    wall0 = time.perf_counter()
    ctx.timer_start()
    s.vv_step(args.dt, args.steps, args.rebin_every)
    ms = ctx.timer_stop()
    barrier()
    wall = time.perf_counter() - wall0
    force_ms, force_launches = s.profile_end()
    kinds = [s.profile_kind(k) for k in range(3)]          # (ms, launches): window scan, list build, list walk
    dom = 2 if kinds[2][1] > 0 else 0                       # the stepping kernel; systems that cannot use a list scan windows
    dom_ms, dom_launches = kinds[dom]
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = allmax(ms)
    pairs1 = allsum(float(s.pair_set_digest()[0]))
    pairs = 0.5 * (pairs0 + pairs1)
    t = ms * 1e-3
    value = pairs * args.steps / t
    atom_steps = N * args.steps / t
    force_ms_per_launch = allmax(dom_ms / max(dom_launches, 1))
    build_ms_per_launch = allmax(kinds[1][0] / max(kinds[1][1], 1))

    # ---- roofline of the dominant kernel (this rank's share of the pairs per launch) ---------------
    nloc, _ = s.local_count()
    pairs_local = pairs * nloc / N
    achieved = FLOPS_PER_PAIR * pairs_local / (force_ms_per_launch * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    rebins = kinds[1][1] if kinds[1][1] > 0 else (args.steps / args.rebin_every if args.rebin_every > 0 else 0)
    force_bytes = BYTES_LIST_STEP if dom == 2 else BYTES_FORCE_EVAL - 16                                # forces only: no e,w
    step_bytes = nloc * (force_bytes + BYTES_VV) + nloc * BYTES_REBIN * rebins / args.steps
    traffic = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload, 1 GPU)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if world == 1 and prof.get("workload") == args.workload and dom == 2:
            traffic = prof["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {
        "bound": "fp64", "kernel": "k_force_list_p" if dom == 2 else "k_force_cells", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
        "frac": achieved / (fp64_peak / 1e12), "traffic": traffic,
        "peak_source": "DFMA chains measured in this run (emdee_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; nominal 37 TFLOP/s",
        "flops_per_pair": FLOPS_PER_PAIR, "pairs_per_launch": pairs_local, "ms_per_launch": force_ms_per_launch,
        "launches_timed": dom_launches, "kernel_share_of_step": dom_ms / ms if ms > 0 else None,
        "kernel_also_does": ("the velocity-Verlet kick and drift of the step (fused into the producer warps; EMDEE_FUSE_VV=0 "
                             "runs them as k_vv and the force kernel alone takes 1.22 ms)") if dom == 2 and world == 1 and os.environ.get("EMDEE_FUSE_VV", "1") != "0" else None,
        "force_kernels_share_of_step": force_ms / ms if ms > 0 else None,
        "list_build": {"kernel": "k_list_build", "ms_per_launch": build_ms_per_launch, "launches_timed": kinds[1][1],
                       "share_of_step": kinds[1][0] / ms if ms > 0 else None},
        "hbm": {"achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "bytes_per_atom_step": step_bytes / max(nloc, 1)},
    }

    # ---- e2e: host buffers in, host buffers out, through the C ABI ---------------------------------
    pos_h = torch.from_numpy(w["pos"].copy()).pin_memory().numpy()
    f_h = torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()
    e_h = torch.empty(N, dtype=torch.float64).pin_memory().numpy()
    w_h = torch.empty(N, dtype=torch.float64).pin_memory().numpy()
    s.set_skin(0.0)

    def e2e_once():
        s.set_positions(pos_h)
        s.bin(args.ndiv)
        s.compute(em.CUTOFF, em.FORCES | em.ENERGIES | em.VIRIALS)
        s.forces(f_h)
        s.energies(e_h)
        s.virials(w_h)

    e2e_once()
    pairs_e2e = allsum(float(s.pair_set_digest()[0]))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_iters):
        e2e_once()
    barrier()
    e2e_t = allmax((time.perf_counter() - t0) / args.e2e_iters)
    e2e = {"value": pairs_e2e / e2e_t, "unit": "pair-interactions/s", "h2d_bytes_per_step": 24 * N,
           "d2h_bytes_per_step": 40 * N, "ms_per_call": e2e_t * 1e3, "atom_evals_per_s": N / e2e_t,
           "call": "set_positions(host) -> bin -> compute_nonbonded(CUTOFF, F|E|V) -> forces/energies/virials(host)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_leg(args, w, budget_s=25.0, steps=3, warmup=1)
        cpu = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "atom_steps_per_s": r["atom_steps_per_s"]}

    if rank == 0:
        line = {
            "metric": "LJ pair-interactions/s", "value": value, "unit": "pair-interactions/s",
            "atom_steps_per_s": atom_steps, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "N": N, "L": L, "cutoff": w["rc"], "switch": w["rs"], "dt": args.dt,
                       "ndiv": args.ndiv, "skin": args.skin,
                       "rebin_every": args.rebin_every if args.rebin_every >= 0 else "adaptive (skin/2 criterion)",
                       "rebins_in_timed_steps": int(rebins),
                       "decomposition": "z-slabs x%d" % world if world > 1 else "single GPU",
                       "l2": "per-step working set %.0f MB exceeds the 126 MB L2" % (N * (BYTES_VV + 48) / 1e6)
                       if N * (BYTES_VV + 48) > 126e6 else "working set fits L2 (consecutive MD steps reuse it by design)",
                       "pairs": pairs, "wall_ms_per_step": wall / args.steps * 1e3},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
