#!/usr/bin/env python
"""bench.py -- throughput of the nonbonded hot path on synthetic FCC Lennard-Jones fluids.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c4|c5]

A "step" is one velocity-Verlet step of the whole fluid: [second half-kick + first half-kick + drift]
kernel, re-binning when an atom has moved more than skin/2 since the last binning (or at a fixed cadence,
--rebin-every), one LJ force evaluation from the pair list.  The timed region holds
exactly K steps issued as ONE emdee_vv_step call, bracketed by a barrier and a device synchronisation, timed
with CUDA events on the library's stream, max over ranks.  Host synchronisations INSIDE that call (they are part
of the measured time): the adaptive re-binning criterion reads one 4-byte word back per step, and every
re-binning reads the brick capacity (and, in a slab decomposition, the slab's atom counts) back once.

metric  : LJ pair-interactions/s = (unique pairs i<j with r2 <= rc2, counted by the audit kernel) x K / t
          ("atom_steps_per_s" is printed beside it: N x K / t) -- BASELINE.json's two headline numbers.
workload: c3 = LJ fluid N=4,000,000 (fcc 100^3), rc=2.5 sigma, rs=2.0, rho*=0.8442 -- the configuration
          BASELINE.json's target is quoted on; it fits one B200, and is strong-scaled over 1/2/4/8 GPUs.
e2e     : the same metric through the C ABI with HOST buffers: positions in pinned host memory ->
          emdee_set_positions (H2D) -> emdee_bin -> emdee_compute_nonbonded_into(F|E|V: the reference's call shape,
          forces / energies / virials written to pinned host memory; chunks of z planes are copied out behind the
          evaluation), every iteration inside the timed region.  Slab ranks move the id window of the atoms they own
          (emdee_set_positions_range, emdee_compute_nonbonded, emdee_get_*_range).
roofline: dominant kernel k_force_list_p (the pair-list stepping kernel, one launch per step); achieved =
          71 flop x pairs per launch / mean launch duration (CUDA events around every launch on the library's
          stream, emdee_profile_begin/end/kind); peak = DFMA throughput measured on this GPU in the same run
          (MEASURED_PEAKS.json holds no FP64 number).  The pair-list build kernel (one launch per re-binning)
          is reported beside it; "traffic" is the DRAM bytes per launch of the ncu capture under profiles/.
cpu_baseline: the OpenMP cell-list oracle on all host cores on a bounded sample of the same workload, plus
          "reference_case": BASELINE configs[0] (N=4000, all pairs) through the C ABI with host arrays next to the
          reference's own CPU path (the serial loop of naively_compute_nonbonded!, one thread).
clocks  : SM clock and throttle reasons sampled through NVML every millisecond inside the timed region.
Other workloads (parity-test configurations, not bench lines of the driver): c1, c2, c5 (FCC fluids) and c4 (the
reference's molecular test system replicated to 1,107,351 atoms, exclusions, 5 LJ classes; lengths in Angstrom).
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
    # config 4: n = replications per dimension of the reference's molecular test system (lengths in A, energies in kJ/mol)
    "c4": (9, 10.0, 9.0, "dibenzo-p-dioxin in water (test/data, 1519 atoms) replicated 9^3 = 1107351 atoms, rc=10 A, 1-2/1-3 exclusions"),
}
MOLECULAR = {"c4"}
GOLDEN_C4 = os.path.join(ROOT, "tests", "golden", "dioxin_water.npz")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--ndiv", type=int, default=1)
    ap.add_argument("--skin", type=float, default=None, help="pair-list skin (default 0.45 sigma; 1.0 A for c4)")
    ap.add_argument("--rebin-every", type=int, default=-1,
                    help="steps between re-binnings; -1: adaptive (re-bin when an atom has moved more than skin/2)")
    ap.add_argument("--dt", type=float, default=None, help="time step (default 0.005 tau; 0.01 = 1 fs for c4)")
    ap.add_argument("--temperature", type=float, default=None, help="kT of the initial velocities (default 1.44 eps; 2.494 kJ/mol for c4)")
    ap.add_argument("--e2e-iters", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the final state (the `parity` block of the line)")
    a = ap.parse_args()
    mol = a.workload in MOLECULAR
    if a.skin is None:
        a.skin = 1.0 if mol else 0.45
    if a.dt is None:
        a.dt = 0.01 if mol else 0.005           # c4: A, amu, kJ/mol -> time unit 0.1 ps
    if a.temperature is None:
        a.temperature = 2.494 if mol else 1.44   # c4: kT at 300 K in kJ/mol
    return a


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled from a
    thread of this process every millisecond (a timed region of 20 steps lasts ~30 ms, which the 20 ms cadence of an
    `nvidia-smi -lms` child process would sample once at best); `mark()` brackets the timed region, and only samples
    taken inside it are reported (all samples, i.e. warm-up + timed region, if the region held fewer than 3).
    Falls back to the nvidia-smi child when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None
        self.samples, self.marks, self.nvml, self.handle, self.stop_flag, self.max_mhz = [], [], None, None, False, None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.device
            if vis:
                ids = [x.strip() for x in vis.split(",") if x.strip()]
                if self.device < len(ids) and ids[self.device].isdigit():
                    index = int(ids[self.device])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def mark(self):
        """Call at the start and at the end of the timed region."""
        self.marks.append(time.perf_counter())

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                why = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((time.perf_counter(), mhz, why))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def _stop_nvml(self):
        self.stop_flag = True
        self.thread.join(timeout=2)
        n = self.nvml
        names = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                 ("hw_power_brake_slowdown", 0x80))        # nvmlClocksEventReason* bits
        inside = self.samples
        region = "warm-up + timed steps"
        if len(self.marks) >= 2:
            sel = [x for x in self.samples if self.marks[0] <= x[0] <= self.marks[-1]]
            if len(sel) >= 3:
                inside, region = sel, "timed steps"
        bits = 0
        for _, _, why in inside:
            bits |= why
        try:
            n.nvmlShutdown()
        except Exception:
            pass
        sm = [x[1] for x in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for name, bit in names if bits & bit), "samples": len(sm), "sampled_over": region,
                "source": "NVML polled in-process every ms"}

    def stop(self):
        if self.nvml is not None:
            return self._stop_nvml()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


def workload(args):
    import emdee_jl_b200 as em

    n, rc, rs, desc = WORKLOADS[args.workload]
    return sample(args, n, rc, rs, desc)


def sample(args, n, rc, rs, desc=""):
    """The workload at size parameter n (fcc cells per dimension, or replications of the molecular fixture)."""
    import emdee_jl_b200 as em

    if args.workload in MOLECULAR:
        m = em.workloads.molecular_system(dict(np.load(GOLDEN_C4)), reps=n)
        pos, L, N = m["positions"], m["L"], m["positions"].shape[0]
        return dict(n=n, rc=rc, rs=rs, desc=desc, pos=pos, L=L, N=N, atoms=m["atoms"], mass=m["masses"], excl=m["excl"],
                    vel=em.workloads.maxwell_velocities(N, args.temperature, m["masses"]))
    pos, L = em.workloads.fcc_lattice(n)
    N = pos.shape[0]
    return dict(n=n, rc=rc, rs=rs, desc=desc, pos=pos, L=L, N=N, atoms=em.workloads.lj_fluid_atoms(N), mass=np.ones(N), excl=None,
                vel=em.workloads.maxwell_velocities(N, args.temperature))


def cpu_leg(args, w, budget_s, steps, warmup):
    """Times the CPU oracle (OpenMP, all host cores) on a bounded sample of the workload: velocity-Verlet
    steps of the same kind of system (same density, cutoff, LJ classes, exclusions) small enough to finish
    in `budget_s`."""
    from oracle import oracle_c

    oracle_c.build()
    if os.environ.get("WORLD_SIZE", "1") != "1" or "OMP_NUM_THREADS" not in os.environ:
        # under torchrun every rank gets OMP_NUM_THREADS=1; this leg runs on rank 0 alone and uses the whole host
        try:
            oracle_c.set_num_threads(len(os.sched_getaffinity(0)), fast=True)
        except AttributeError:
            oracle_c.set_num_threads(os.cpu_count() or 1, fast=True)
    cores = oracle_c.num_threads(fast=True)
    mol = args.workload in MOLECULAR
    # probe the evaluation rate on a small sample, then pick the largest sample that fits the budget
    p0 = sample(args, 2 if mol else 16, w["rc"], w["rs"])
    oracle_c.cutoff_cells(p0["pos"], p0["L"], w["rc"], w["rs"], p0["atoms"], ndiv=args.ndiv, excl=p0["excl"], bitmask=1, fast=True)
    t0 = time.perf_counter()
    oracle_c.cutoff_cells(p0["pos"], p0["L"], w["rc"], w["rs"], p0["atoms"], ndiv=args.ndiv, excl=p0["excl"], bitmask=1, fast=True)
    rate = p0["N"] / (time.perf_counter() - t0)          # atom-evaluations/s, O(N)
    per_n3 = p0["N"] / p0["n"] ** 3                       # atoms per unit of n^3
    n, nmin = w["n"], (2 if mol else 8)
    while n > nmin and per_n3 * n ** 3 * (steps + warmup + 1) / rate > budget_s:
        n = max(nmin, int(n * 0.8))
    q = sample(args, n, w["rc"], w["rs"])
    N, L, at, mass, excl = q["N"], q["L"], q["atoms"], q["mass"], q["excl"]
    r = oracle_c.cutoff_cells(q["pos"], L, w["rc"], w["rs"], at, ndiv=args.ndiv, excl=excl, bitmask=1, fast=True)
    p, v, f = q["pos"], q["vel"], r["forces"]
    if warmup:
        p, v, f = oracle_c.vv_steps(p, v, f, mass, L, w["rc"], w["rs"], at, args.dt, warmup, ndiv=args.ndiv, excl=excl, fast=True)
    t0 = time.perf_counter()
    p, v, f = oracle_c.vv_steps(p, v, f, mass, L, w["rc"], w["rs"], at, args.dt, steps, ndiv=args.ndiv, excl=excl, fast=True)
    dt = time.perf_counter() - t0
    npairs = oracle_c.cutoff_cells(p, L, w["rc"], w["rs"], at, ndiv=args.ndiv, excl=excl, bitmask=1, fast=True)["npairs"]
    kind = ("the molecular fixture replicated %d^3" % n) if mol else "an fcc LJ fluid"
    return dict(value=npairs * steps / dt, unit="pair-interactions/s", cores=cores, kind="port",
                sample="%d velocity-Verlet steps of N=%d atoms of %s (same density, rc, rs, dt%s; full-neighbour "
                       "OpenMP cell-list oracle, -O3 AVX2+FMA)" % (steps, N, kind, ", exclusions" if mol else ""),
                atom_steps_per_s=N * steps / dt, ms_per_step=dt / steps * 1e3, N=N)


def reference_case(em, ctx):
    """BASELINE configs[0] beside the headline: N = 4000 (fcc 10^3), rc = 2.5, single-point energies / forces / virials
    with the REFERENCE'S OWN semantics and call shape -- `compute_nonbonded!(f, e, w, positions, L, tiles, model, atoms,
    Val(7))` over all N(N-1)/2 minimum-image pairs (src/nonbonded.jl:109-120) -- through the C ABI with host arrays
    (upload, k_force_tiles, download), next to the reference's own CPU path, the serial loop of
    `naively_compute_nonbonded!` (src/nonbonded.jl:122-155), restated in oracle/ and timed on one host thread."""
    from oracle import oracle_c

    pos, L = em.workloads.fcc_lattice(10)
    N = pos.shape[0]
    atoms = em.workloads.lj_fluid_atoms(N)
    model = em.LennardJonesModel(2.5, 2.0)
    tiles = em.nonbonded_computation_tiles(N)
    f = np.zeros((N, 3)); e = np.zeros(N); w = np.zeros(N)
    s = em.NonbondedSystem(N, L, ctx)
    s.set_model(model)
    s.set_atoms(atoms)
    s.set_tiles(tiles)

    def once():
        s.set_positions(pos)
        s.compute(em.ALLPAIRS_REFERENCE, 7)
        s.forces(f); s.energies(e); s.virials(w)

    once()
    t0 = time.perf_counter()
    for _ in range(5):
        once()
    gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    s.close()
    om = oracle_c.lj_model(2.5, 2.0)
    oracle_c.naive_allpairs(pos, L, om, atoms)
    t0 = time.perf_counter()
    fr, er, wr = oracle_c.naive_allpairs(pos, L, om, atoms)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    npairs = N * (N - 1) // 2
    frms = float(np.sqrt((fr ** 2).sum(axis=1).mean()))
    return {"workload": "LJ fluid N=4000 (fcc 10^3), rc=2.5: compute_nonbonded!(..., Val(7)) over all %d pairs, host arrays in and out" % npairs,
            "b200_ms_per_call": gpu_ms, "b200_pairs_per_s": npairs / (gpu_ms * 1e-3),
            "cpu_ms_per_call": cpu_ms, "cpu_pairs_per_s": npairs / (cpu_ms * 1e-3), "cpu_threads": 1,
            "cpu_kind": "port of naively_compute_nonbonded! (serial loop, FP64)",
            "max_force_diff_over_frms": float(np.abs(f - fr).max() / frms),
            "energy_rel_diff": float(abs(e.sum() - er.sum()) / abs(er.sum()))}


def common_config(args, w, N, L):
    """The keys both arms print under `config` (identical for the same command line): the workload and how the timed
    steps relate to the 126 MB L2."""
    return {"workload": w["desc"], "N": int(N), "L": float(L), "cutoff": w["rc"], "switch": w["rs"], "dt": args.dt, "ndiv": args.ndiv,
            "l2": "per-step working set %.0f MB exceeds the 126 MB L2" % (N * (BYTES_VV + 48) / 1e6)
            if N * (BYTES_VV + 48) > 126e6 else "working set fits L2 (consecutive MD steps reuse it by design)"}


def parity_block(args, em, s, w, world, rank, allsum_arr):
    """Parity of THIS run's final state against the CPU oracle, printed in the bench line so that the driver's own
    1/2/4/8-GPU runs carry it ("parity under decomposition"): positions after the timed steps are gathered
    in id order, and (i) the forces the stepping path left behind, (ii) a single-point F/E/W evaluation, (iii) the pair-set
    digest (count, sum and xor of the pair hashes over all ranks) are compared with the oracle's on those positions.  The
    oracle is the checker only; nothing here is timed."""
    N, L = w["N"], w["L"]
    s.synchronize()
    pos = allsum_arr(s.positions())
    f_step = allsum_arr(s.forces())
    s.compute(em.CUTOFF, em.FORCES | em.ENERGIES | em.VIRIALS)
    f_sp = allsum_arr(s.forces())
    e_sp = allsum_arr(s.energies())
    w_sp = allsum_arr(s.virials())
    dig = np.asarray(s.pair_set_digest(), dtype=np.uint64)
    halves = np.array([int(dig[0]), int(dig[1]) & 0xFFFFFFFF, int(dig[1]) >> 32], dtype=np.int64)     # wrap-around sum in two halves
    halves = allsum_arr(halves)
    xor = int(dig[2])
    if world > 1:
        import torch
        import torch.distributed as dist

        parts = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(parts, torch.tensor([xor & 0xFFFFFFFF, xor >> 32], dtype=torch.int64, device="cuda"))
        xor = 0
        for t in parts:
            lo, hi = (int(x) for x in t.cpu())
            xor ^= lo | (hi << 32)
    s.compute(em.CUTOFF, em.FORCES)          # leave the system as the stepping path expects it
    if rank != 0:
        return None
    from oracle import oracle_c

    oracle_c.build()
    try:
        oracle_c.set_num_threads(len(os.sched_getaffinity(0)), fast=True)
    except AttributeError:
        oracle_c.set_num_threads(os.cpu_count() or 1, fast=True)
    ref = oracle_c.cutoff_cells(pos, L, w["rc"], w["rs"], w["atoms"], ndiv=args.ndiv, excl=w["excl"], fast=True)
    frms = float(np.sqrt((ref["forces"] ** 2).sum(axis=1).mean()))
    got = [int(halves[0]), (int(halves[1]) + (int(halves[2]) << 32)) % (1 << 64), xor]
    want = [int(x) for x in ref["digest"]]
    out = {"checked_against": "CPU oracle (oracle/, OpenMP) on this run's final positions, N=%d, %d rank(s)" % (N, world),
           "pair_digest_equal": got == want, "pairs": [got[0], int(ref["npairs"])],
           "force_err_over_frms_stepping": float(np.abs(f_step - ref["forces"]).max() / frms),
           "force_err_over_frms_single_point": float(np.abs(f_sp - ref["forces"]).max() / frms),
           "E_rel_err": float(abs(e_sp.sum() - ref["E"]) / abs(ref["E"])), "W_rel_err": float(abs(w_sp.sum() - ref["W"]) / abs(ref["W"])),
           "tolerances": {"pairs": "bit-exact", "E,W": 1e-10, "forces": "1e-9 F_rms"}}
    out["ok"] = bool(out["pair_digest_equal"] and out["force_err_over_frms_stepping"] <= 1e-9 and out["force_err_over_frms_single_point"] <= 1e-9
                     and out["E_rel_err"] <= 1e-10 and out["W_rel_err"] <= 1e-10)
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Julia
    (not installed; nothing under /root/reference compiles with gcc), so oracle/_ref does not exist and
    this arm times the oracle port on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(zip(("n", "rc", "rs", "desc"), WORKLOADS[args.workload]))
    r = cpu_leg(args, w, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    import emdee_jl_b200 as em

    if args.workload in MOLECULAR:
        fx = np.load(GOLDEN_C4)
        Nfull, Lfull = int(fx["positions"].shape[0]) * w["n"] ** 3, float(fx["box"]) * w["n"]
    else:
        Nfull, Lfull = 4 * w["n"] ** 3, em.workloads.fcc_box(w["n"])[1]
    line = {
        "impl": "reference", "metric": "LJ pair-interactions/s", "value": r["value"], "unit": "pair-interactions/s",
        "atom_steps_per_s": r["atom_steps_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": common_config(args, w, Nfull, Lfull),
        "run": {"sample": r["sample"], "sample_N": r["N"]},
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

    # rebin_every = -1 (default): one GPU re-bins on the step on which an atom has moved more than skin/2 (a 4-byte read-back per
    # step); slab runs choose the interval at every re-binning from the largest displacement of the interval that ended there
    # (max over ranks, read back with the re-binning's own synchronisation), so their steps need no read-back at all
    w = workload(args)
    N, L = w["N"], w["L"]
    s = em.NonbondedSystem(N, L, ctx)
    s.set_model(em.LennardJonesModel(w["rc"], w["rs"]))
    s.set_atoms(w["atoms"])
    s.set_positions(w["pos"])
    s.set_velocities(w["vel"])
    s.set_masses(w["mass"])
    if w["excl"] is not None:
        s.set_exclusions(*w["excl"])
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

    def allsum_arr(a):
        if world == 1:
            return a
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fp64_peak = ctx.measure_fp64_peak()
    cfg = s.step_config()
    step_kernel = ("k_force_list_p" if cfg["persistent"] else "k_force_list") if cfg["pair_list"] else "k_force_cells"

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
    sc0 = s.step_counters()
    s.profile_begin()
    sampler.mark()
    wall0 = time.perf_counter()
    ctx.timer_start()
    s.vv_step(args.dt, args.steps, args.rebin_every)
    ms = ctx.timer_stop()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark()
    force_ms, force_launches = s.profile_end()
    sc1 = s.step_counters()
    list_modes = {k: sc1[k] - sc0[k] for k in ("walk", "prune", "replay")}     # stepping launches by list mode (two-level list)
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
    # list walk: 2 B per entry inside rc + skin (the mean over atoms) + 54 B recipe + 24 B positions + 24 B forces
    list_bytes = 2.0 * N / L ** 3 * 4.18879 * (w["rc"] + args.skin) ** 3 + (BYTES_LIST_STEP - 176)
    force_bytes = list_bytes if dom == 2 else BYTES_FORCE_EVAL - 16                                # forces only: no e,w
    step_bytes = nloc * (force_bytes + BYTES_VV) + nloc * BYTES_REBIN * rebins / args.steps
    traffic = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload, 1 GPU)
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        prof = json.load(open(tp if os.path.exists(tp) else os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if world == 1 and prof.get("workload") == args.workload and dom == 2:
            traffic = prof["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {
        "bound": "fp64", "kernel": step_kernel if dom == 2 else "k_force_cells", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
        "frac": achieved / (fp64_peak / 1e12), "traffic": traffic,
        "peak_source": "DFMA chains measured in this run (emdee_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry; nominal 37 TFLOP/s",
        "flops_per_pair": FLOPS_PER_PAIR, "pairs_per_launch": pairs_local, "ms_per_launch": force_ms_per_launch,
        "launches_timed": dom_launches, "kernel_share_of_step": dom_ms / ms if ms > 0 else None,
        "kernel_also_does": ("the velocity-Verlet kick and drift of the step (fused into the producer warps; EMDEE_FUSE_VV=0 "
                             "runs them as k_vv)") if dom == 2 and cfg["fused_vv"] else None,
        "launches_by_list_mode": list_modes,
        "force_kernels_share_of_step": force_ms / ms if ms > 0 else None,
        "list_build": {"kernel": "k_list_build", "ms_per_launch": build_ms_per_launch, "launches_timed": kinds[1][1],
                       "share_of_step": kinds[1][0] / ms if ms > 0 else None},
        "hbm": {"achieved": step_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                "bytes_per_atom_step": step_bytes / max(nloc, 1)},
    }

    # ---- parity of this run's final state against the CPU oracle (outside every timed region) ------
    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(args, em, s, w, world, rank, allsum_arr)
        except Exception as ex:                      # never lose the headline line over the check; a failure is visible in the line
            parity = {"ok": False, "error": repr(ex)}

    # ---- e2e: host buffers in, host buffers out, through the C ABI ---------------------------------
    # One GPU: the whole arrays.  Slab ranks: every rank moves the window of the id-ordered host arrays that covers the atoms it
    # owns (emdee_get_local_id_range; for these lattices ids run with z, so a window is about N/ranks rows) and receives the
    # rows of those atoms -- per call the ranks together move about what one GPU moves alone.  The windows are those of the
    # decomposition of the uploaded positions themselves (one untimed full-array call first).
    s.set_skin(0.0)
    s.set_positions(w["pos"])
    s.bin(args.ndiv)
    id0, cnt = s.local_id_range() if world > 1 else (0, N)
    pos_h = torch.from_numpy(w["pos"].copy()).pin_memory().numpy()           # full id-ordered host arrays, pinned; a slab rank
    f_h = torch.empty((N, 3), dtype=torch.float64).pin_memory().numpy()      # touches only its window's rows of them
    e_h = torch.empty(N, dtype=torch.float64).pin_memory().numpy()
    w_h = torch.empty(N, dtype=torch.float64).pin_memory().numpy()

    def e2e_once():
        if world > 1:
            s.set_positions_range(id0, cnt, pos_h)
        else:
            s.set_positions(pos_h)
        s.bin(args.ndiv)
        if world > 1:
            s.compute(em.CUTOFF, em.FORCES | em.ENERGIES | em.VIRIALS)
            s.forces_range(id0, cnt, f_h)
            s.energies_range(id0, cnt, e_h)
            s.virials_range(id0, cnt, w_h)
        else:     # the reference's own call shape: compute_nonbonded!(forces, energies, virials, ...) fills three host arrays
            s.compute_into(em.CUTOFF, em.FORCES | em.ENERGIES | em.VIRIALS, f_h, e_h, w_h)

    e2e_once()
    pairs_e2e = allsum(float(s.pair_set_digest()[0]))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_iters):
        e2e_once()
    barrier()
    e2e_t = allmax((time.perf_counter() - t0) / args.e2e_iters)
    rows = allsum(float(cnt))
    e2e = {"value": pairs_e2e / e2e_t, "unit": "pair-interactions/s", "h2d_bytes_per_step": int(24 * rows),
           "d2h_bytes_per_step": int(40 * rows), "ms_per_call": e2e_t * 1e3, "atom_evals_per_s": N / e2e_t,
           "call": ("set_positions(host) -> bin -> compute_nonbonded_into(CUTOFF, F|E|V, forces, energies, virials: host)" if world == 1 else
                    "per rank, on the cyclic id window of the atoms it owns: set_positions_range(host) -> bin -> compute_nonbonded(CUTOFF, F|E|V) -> "
                    "forces/energies/virials_range(host); bytes are summed over ranks")}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_leg(args, w, budget_s=25.0, steps=3, warmup=1)
        cpu = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "atom_steps_per_s": r["atom_steps_per_s"]}
        try:
            cpu["reference_case"] = reference_case(em, ctx)
        except Exception as ex:                      # never lose the headline line over the side measurement
            cpu["reference_case"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "LJ pair-interactions/s", "value": value, "unit": "pair-interactions/s",
            "atom_steps_per_s": atom_steps, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": common_config(args, w, N, L),
            "run": {"skin": args.skin, "brick_cells": list(cfg["brick"]), "brick_capacity": cfg["brick_capacity"],
                    "rebin_every": args.rebin_every if args.rebin_every >= 0 else ("adaptive (skin/2 criterion)" if world == 1 else
                                   "adaptive (interval chosen at every re-binning from the last interval's largest displacement)"),
                    "rebins_in_timed_steps": int(rebins),
                    "decomposition": "z-slabs x%d" % world if world > 1 else "single GPU",
                    "pairs": pairs, "wall_ms_per_step": wall / args.steps * 1e3},
            "parity": parity,
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
        try:
            run_b200(a)
        except BaseException:
            # one failing rank must not leave the others waiting in a collective until the launcher's timeout
            import traceback

            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)
