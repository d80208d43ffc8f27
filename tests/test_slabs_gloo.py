"""world_size-2 gloo test (CPU) of the slab decomposition's host logic: plane ownership, the halo plan
and its message order, and that owned + ghost atoms reproduce the global pair set exactly once."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ndiv, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import emdee_jl_b200 as em
    from emdee_jl_b200 import slabs
    from oracle import oracle_c as oc

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    pos, L = em.workloads.fcc_lattice(10)
    N = pos.shape[0]
    rc = 2.5
    M = oc.cells_per_dimension(L, rc, ndiv)
    R = ndiv
    assert slabs.check(world, M, R)
    idx = oc.cell_index(pos, L, M) - 1
    z = idx // (M * M)
    z0, z1 = slabs.plane_range(rank, world, M)
    own = np.nonzero((z >= z0) & (z < z1))[0]
    plan = slabs.halo_plan(rank, world, M, R)
    lower, upper = slabs.neighbours(rank, world)

    def ids_in(planes):
        # contiguous in the (cell, id) order: planes are consecutive z values
        sel = np.nonzero(np.isin(z, planes))[0]
        return sel[np.lexsort((sel, idx[sel]))].astype(np.int64)

    send_lo, send_hi = ids_in(plan["send_to_lower"]), ids_in(plan["send_to_upper"])
    # counts first, then payloads, in the library's message order
    cnt_hi, cnt_lo = torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64)
    ops = [dist.P2POp(dist.isend, torch.tensor([send_lo.size]), lower), dist.P2POp(dist.isend, torch.tensor([send_hi.size]), upper),
           dist.P2POp(dist.irecv, cnt_hi, upper), dist.P2POp(dist.irecv, cnt_lo, lower)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    ghost_hi, ghost_lo = torch.zeros(int(cnt_hi), dtype=torch.int64), torch.zeros(int(cnt_lo), dtype=torch.int64)
    ops = [dist.P2POp(dist.isend, torch.from_numpy(send_lo), lower), dist.P2POp(dist.isend, torch.from_numpy(send_hi), upper),
           dist.P2POp(dist.irecv, ghost_hi, upper), dist.P2POp(dist.irecv, ghost_lo, lower)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    ok = np.array_equal(ghost_lo.numpy(), ids_in(plan["recv_lower_ghosts"]))
    ok &= np.array_equal(ghost_hi.numpy(), ids_in(plan["recv_upper_ghosts"]))
    # pairs seen by this rank: owned i against owned + ghost j, counted when id_i < id_j
    local = np.concatenate([ghost_lo.numpy(), own, ghost_hi.numpy()])
    pairs = oc.pair_set_brute(pos, L, rc * rc)[0]
    local_set = set(local.tolist())
    own_set = set(own.tolist())
    mine = [(i, j) for i, j in pairs.tolist() if (i in own_set and j in local_set) or (j in own_set and i in local_set)]
    # every global pair touching an owned atom must be visible locally (ghost layer is thick enough)
    touching = [(i, j) for i, j in pairs.tolist() if i in own_set or j in own_set]
    ok &= len(mine) == len(touching)
    counted = sum(1 for i, j in mine if i in own_set)        # rule: the rank owning the smaller id counts the pair
    tot = torch.tensor([counted], dtype=torch.int64)
    dist.all_reduce(tot)
    nown = torch.tensor([own.size], dtype=torch.int64)
    dist.all_reduce(nown)
    ok &= int(tot) == pairs.shape[0] and int(nown) == N
    # ---- migration at a re-binning: the atoms drift (a fraction of a plane), leavers go to the ring neighbours ----
    rng = np.random.default_rng(5)
    moved = pos + rng.normal(scale=0.25, size=pos.shape)        # same numbers on every rank
    znew = (oc.cell_index(moved, L, M) - 1) // (M * M)
    tgt = np.array(slabs.migration_targets(znew[own], rank, world, M))
    go_lo, go_hi = own[tgt == -1].astype(np.int64), own[tgt == 1].astype(np.int64)
    n_hi, n_lo = torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64)
    ops = [dist.P2POp(dist.isend, torch.tensor([go_lo.size]), lower), dist.P2POp(dist.isend, torch.tensor([go_hi.size]), upper),
           dist.P2POp(dist.irecv, n_hi, upper), dist.P2POp(dist.irecv, n_lo, lower)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    in_hi, in_lo = torch.zeros(int(n_hi), dtype=torch.int64), torch.zeros(int(n_lo), dtype=torch.int64)
    ops = [dist.P2POp(dist.isend, torch.from_numpy(go_lo), lower), dist.P2POp(dist.isend, torch.from_numpy(go_hi), upper),
           dist.P2POp(dist.irecv, in_hi, upper), dist.P2POp(dist.irecv, in_lo, lower)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    new_own = np.sort(np.concatenate([own[tgt == 0], in_lo.numpy(), in_hi.numpy()]))
    expect = np.nonzero((znew >= z0) & (znew < z1))[0]
    ok &= np.array_equal(new_own, expect) and (go_lo.size + go_hi.size) > 0
    q.put((rank, bool(ok), int(tot), pairs.shape[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ndiv", [(2, 1), (2, 2), (3, 1)])
def test_slab_plan_world2(world, ndiv):
    """world 2: both ring neighbours are the same peer (message order matters); world 3: distinct neighbours."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ndiv, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res


def test_plane_ranges():
    sys.path.insert(0, ROOT)
    from emdee_jl_b200 import slabs

    for M, G in ((67, 8), (6, 2), (134, 8), (26, 4), (13, 2)):
        planes = []
        for r in range(G):
            z0, z1 = slabs.plane_range(r, G, M)
            assert z1 - z0 in (M // G, M // G + 1)
            planes += list(range(z0, z1))
        assert planes == list(range(M))
        assert all(slabs.owner_of_plane(z, G, M) == r for r in range(G) for z in range(*slabs.plane_range(r, G, M)))
    assert slabs.neighbours(0, 8) == (7, 1) and slabs.neighbours(1, 2) == (0, 0)
    assert not slabs.check(8, 6, 1) and slabs.check(8, 67, 1) and not slabs.check(2, 4, 2)
    p = slabs.halo_plan(0, 2, 6, 1)
    assert p["send_to_lower"] == [0] and p["send_to_upper"] == [2] and p["recv_lower_ghosts"] == [5] and p["recv_upper_ghosts"] == [3]
