"""slab_worker.py -- multi-GPU parity check, one process per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/slab_worker.py
Every rank builds the same FCC fluid, the library decomposes it into z slabs (NCCL halo exchange and
migration), and the union of the ranks' results is compared with the CPU oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import emdee_jl_b200 as em
    from oracle import oracle_c as oc

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = em.Context(local)
    ids = [em.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    n = int(os.environ.get("SLAB_N", "16"))
    ndiv = int(os.environ.get("SLAB_NDIV", "1"))
    pos, L = em.workloads.fcc_lattice(n)
    N = pos.shape[0]
    atoms = em.workloads.lj_fluid_atoms(N)
    vel = em.workloads.maxwell_velocities(N, 1.44)
    res = {}
    s = em.NonbondedSystem(N, L, ctx)
    s.set_model(em.LennardJonesModel(2.5, 2.0))
    s.set_atoms(atoms)
    s.set_positions(pos)
    s.set_velocities(vel)
    s.bin(ndiv)
    nloc, nghost = s.local_count()
    res["nlocal_sum"] = int(allsum(np.array([nloc], dtype=np.int64))[0])
    s.compute(em.CUTOFF, 7)
    f = allsum(s.forces()); e = allsum(s.energies()); w = allsum(s.virials())
    dig = s.pair_set_digest()
    cnt = allsum(np.array([int(dig[0])], dtype=np.int64))[0]
    hs = allsum(np.array([int(dig[1]) & 0x7FFFFFFF, int(dig[1]) >> 31], dtype=np.int64))     # wrap-around sum in two halves
    ref = oc.cutoff_cells(pos, L, 2.5, 2.0, atoms, ndiv=ndiv, fast=True)
    frms = np.sqrt((ref["forces"] ** 2).sum(1).mean())
    res["force_err"] = float(np.abs(f - ref["forces"]).max() / frms)
    res["E_err"] = float(abs(e.sum() - ref["E"]) / abs(ref["E"]))
    res["W_err"] = float(abs(w.sum() - ref["W"]) / abs(ref["W"]))
    res["pairs"] = [int(cnt), int(ref["npairs"])]
    ref_sum = int(ref["digest"][1])
    got_sum = (int(hs[0]) + (int(hs[1]) << 31)) % (1 << 64)
    res["digest_sum_ok"] = got_sum == ref_sum
    idx = allsum(s.cell_index().astype(np.int64))
    M = oc.cells_per_dimension(L, 2.5, ndiv)
    res["cell_index_ok"] = bool(np.array_equal(idx, oc.cell_index(pos, L, M)))
    # velocity-Verlet with re-binning (migration across slabs) against the oracle
    s.set_skin(0.5)
    s.bin(ndiv)
    s.compute(em.CUTOFF, em.FORCES)
    res["step_config"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in s.step_config().items()}
    ids0 = set(s.local_ids().tolist())
    # 23 + 17 steps in two calls: the first step of a call exchanges the halo by ncclSend/ncclRecv, the following ones get it
    # through peer memory (fused integrator), every fifth step re-bins (migration); 40 steps in all
    nsteps, dt = 40, 0.005
    s.vv_step(dt, 23, rebin_every=5)
    s.vv_step(dt, nsteps - 23, rebin_every=-1)        # interval chosen by the library at every re-binning
    s.synchronize()
    ids1 = set(s.local_ids().tolist())
    res["migrated_out_of_this_rank0_slab"] = len(ids0 - ids1)
    p1 = allsum(s.positions()); v1 = allsum(s.velocities()); f1 = allsum(s.forces())
    f0 = oc.cutoff_cells(pos, L, 2.5, 2.0, atoms, ndiv=ndiv, fast=True)["forces"]
    po, vo, fo = oc.vv_steps(pos, vel, f0, np.ones(N), L, 2.5, 2.0, atoms, dt, nsteps, ndiv=ndiv, fast=True)
    res["vv_pos_err"] = float(np.abs(p1 - po).max())
    res["vv_vel_err"] = float(np.abs(v1 - vo).max())
    res["vv_force_err"] = float(np.abs(f1 - fo).max() / frms)
    # the same bar as a single point, free of trajectory divergence: the forces the stepping path left behind against the
    # oracle evaluated at the GPUs' own final positions, and the evaluated pair count
    ref1 = oc.cutoff_cells(p1, L, 2.5, 2.0, atoms, ndiv=ndiv, fast=True)
    res["vv_force_err_same_positions"] = float(np.abs(f1 - ref1["forces"]).max() / np.sqrt((ref1["forces"] ** 2).sum(1).mean()))
    npl = s.list_pair_count()
    res["vv_list_pairs"] = [int(allsum(np.array([npl], dtype=np.int64))[0]), int(ref1["npairs"])]
    nloc2, _ = s.local_count()
    res["nlocal_sum_after"] = int(allsum(np.array([nloc2], dtype=np.int64))[0])
    # the same single point through the full-array calls (set_positions / forces / energies) ...
    s.set_positions(p1)
    s.bin(ndiv)
    s.compute(em.CUTOFF, 7)
    ff = allsum(s.forces()); ef = allsum(s.energies())
    res["full_force_err"] = float(np.abs(ff - ref1["forces"]).max() / np.sqrt((ref1["forces"] ** 2).sum(1).mean()))
    res["full_E_err"] = float(abs(ef.sum() - ref1["E"]) / abs(ref1["E"]))
    # ... and through the id windows: every rank uploads only the rows of the atoms it owns and downloads those rows
    lo, cnt = s.local_id_range()
    res["id_window_rows_sum"] = int(allsum(np.array([cnt], dtype=np.int64))[0])
    s.set_positions_range(lo, cnt, p1)
    s.bin(ndiv)
    s.compute(em.CUTOFF, 7)
    lo, cnt = s.local_id_range()         # ownership is that of the new binning
    fw = np.zeros((N, 3)); ew = np.zeros(N)
    s.forces_range(lo, cnt, fw)          # only the window's rows of the full arrays are written
    s.energies_range(lo, cnt, ew)
    res["window"] = [int(lo), int(cnt)]
    fw = allsum(fw); ew = allsum(ew)
    bad = np.nonzero(np.abs(fw - ref1["forces"]).max(axis=1) > 1e-6)[0]
    res["window_bad_rows"] = [int(bad.size)] + bad[:8].tolist()
    res["window_force_err"] = float(np.abs(fw - ref1["forces"]).max() / np.sqrt((ref1["forces"] ** 2).sum(1).mean()))
    res["window_E_err"] = float(abs(ew.sum() - ref1["E"]) / abs(ref1["E"]))
    ok = (res["nlocal_sum"] == N and res["nlocal_sum_after"] == N and res["force_err"] <= 1e-9 and res["E_err"] <= 1e-10
          and res["W_err"] <= 1e-10 and res["pairs"][0] == res["pairs"][1] and res["digest_sum_ok"] and res["cell_index_ok"]
          and res["vv_pos_err"] <= 1e-10 and res["vv_vel_err"] <= 1e-9 and res["vv_force_err"] <= 1e-8
          and res["vv_force_err_same_positions"] <= 1e-9 and res["window_force_err"] <= 1e-9 and res["window_E_err"] <= 1e-10
          and res["full_force_err"] <= 1e-9 and res["full_E_err"] <= 1e-10
          and res["id_window_rows_sum"] < 1.25 * N
          and abs(res["vv_list_pairs"][0] - res["vv_list_pairs"][1]) <= world)      # every rank halves its own ordered-pair count
    res["ok"] = bool(ok)
    res["world"] = world
    if rank == 0:
        print("SLAB_RESULT " + json.dumps(res), flush=True)
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:      # a failing rank must not leave the others waiting in a collective
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
