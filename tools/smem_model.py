"""smem_model.py -- CPU model of the shared-memory wavefronts of the stepping kernel's gathers (k_force_list_p).

ncu says the kernel is co-limited by the shared-memory crossbar (profiles/README.md): 217 M wavefronts per launch
= 1736 per warp task at N = 4 M, 6.1 wavefronts per LDS instruction.  This script rebuilds, on the CPU, the exact
index streams a warp task issues (brick staging order, per-lane pair lists, FP16 pre-cull survivors, LIFO drain) for
a perturbed FCC fluid of the bench's density and counts bank-conflict wavefronts under the hardware's rules
(32 banks x 4 B; LDS.64 is served per half-warp, LDS.128 per quarter-warp; lanes reading the same word share it),
first for the layout in the tree -- to calibrate the model against ncu -- then for candidate layouts.

    python tools/smem_model.py [n_fcc_cells=24] [bricks_sampled=40]
"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import emdee_jl_b200 as em  # noqa: E402  (host-side workload generator only; nothing is computed on a GPU here)

RC, SKIN = 2.5, 0.45
BX, BY, BZ = 4, 2, 2


def wavefronts(byte_addr, width, active=None):
    """Wavefronts of one warp-wide shared-memory load: `width` bytes per lane at byte_addr[lane] (aligned to width).
    Lanes are served in groups of 128 B / width... i.e. 16 lanes for 8 B, 8 lanes for 16 B, 32 for 4 B; within a group
    the cost is the largest number of DISTINCT 4-byte words that fall into one bank."""
    lanes_per_group = {4: 32, 8: 16, 16: 8}[width]
    total = 0
    for g0 in range(0, 32, lanes_per_group):
        words = set()
        for lane in range(g0, g0 + lanes_per_group):
            if active is not None and not active[lane]:
                continue
            w0 = int(byte_addr[lane]) // 4
            for k in range(width // 4):
                words.add(w0 + k)
        if not words:
            continue
        per_bank = np.bincount(np.fromiter((w % 32 for w in words), dtype=np.int64), minlength=32)
        total += int(per_bank.max())
    return total


def build_bricks(pos, L):
    M = int(np.floor(L / (RC + SKIN)))
    s = pos / L
    s -= np.floor(s)
    v = np.minimum((s * M).astype(np.int64), M - 1)
    cell = v[:, 0] + M * (v[:, 1] + M * v[:, 2])
    order = np.lexsort((np.arange(len(pos)), cell))               # (cell, id) order
    start = np.searchsorted(cell[order], np.arange(M ** 3 + 1))
    return M, order, start


def brick_streams(pos, L, M, order, start, hx0, hy0, hz0):
    """Staged order of one brick (rows z, y; cells x fastest; ids inside a cell), its home atoms in home-row order,
    and for every home atom the sorted staged indices (1-based, 0 = dummy) inside rc + skin and inside rc."""
    staged, home_pos_in_staged = [], []
    for cz in range(hz0 - 1, hz0 + BZ + 1):
        for cy in range(hy0 - 1, hy0 + BY + 1):
            for cx in range(hx0 - 1, hx0 + BX + 1):
                c = (cx % M) + M * ((cy % M) + M * (cz % M))
                ids = order[start[c]:start[c + 1]]
                is_home = hx0 <= cx < hx0 + BX and hy0 <= cy < hy0 + BY and hz0 <= cz < hz0 + BZ
                for a in ids:
                    if is_home:
                        home_pos_in_staged.append(len(staged))
                    staged.append(a)
    staged = np.array(staged)
    p = pos[staged]
    home = np.array(home_pos_in_staged)
    lists, inside = [], []
    for h in home:
        d = p - p[h]
        d -= L * np.rint(d / L)
        r2 = (d * d).sum(axis=1)
        r2[h] = 1e9
        idx = np.nonzero(r2 <= (RC + SKIN) ** 2 * 1.01)[0]        # FP16 threshold: ~1 % head-room
        lists.append(idx + 1)
        inside.append(r2[idx] <= RC * RC * 1.03)                    # survivors of the walk's FP16 test
    return len(staged), home, lists, inside


def round_robin(x, lane, nclass, pad=False):
    """Order the staged indices x so that the k-th one falls into bank class (lane + k) mod nclass where possible:
    take the next entry of the wanted class; if that class is exhausted either emit a dummy (pad) or take from the
    next non-empty class in cyclic order (no padding, some conflicts remain)."""
    buckets = [[] for _ in range(nclass)]
    for j in x:
        buckets[int(j) % nclass].append(int(j))
    for b in buckets:
        b.reverse()
    out, left, k = [], len(x), 0
    while left:
        c = (lane + k) % nclass
        if buckets[c]:
            out.append(buckets[c].pop()); left -= 1
        elif pad:
            out.append(0)
        else:
            for d in range(1, nclass):
                if buckets[(c + d) % nclass]:
                    out.append(buckets[(c + d) % nclass].pop()); left -= 1
                    break
        k += 1
    return np.array(out, dtype=np.int64)


def task_cost(lists, inside, layout):
    if layout.get("list_order"):
        nclass, pad = layout["list_order"]
        new_lists, new_inside = [], []
        for lane, (x, m) in enumerate(zip(lists, inside)):
            keep = set(int(j) for j in x[m])
            y = round_robin(x, lane, nclass, pad)
            new_lists.append(y)
            new_inside.append(np.array([int(j) in keep for j in y], dtype=bool))
        lists, inside = new_lists, new_inside
    if layout.get("drain_order"):
        # class-aware pops: the survivors of a lane are popped so that row r prefers bank class (lane + r) mod nclass
        nclass = layout["drain_order"]
        stacks = [round_robin(x[m], lane, nclass)[::-1] for lane, (x, m) in enumerate(zip(lists, inside))]   # [::-1]: the model pops from the end
        inside = [np.ones(len(st), dtype=bool) for st in stacks]
        return _task_cost(lists, stacks, inside, layout)
    return _task_cost(lists, None, inside, layout)


def _task_cost(lists, stacks_override, inside, layout):
    """Wavefronts of one warp task (32 lanes) under `layout` -> dict(test=, drain=, stack=)."""
    nmax = max(len(x) for x in lists)
    nmax8 = (nmax + 7) // 8 * 8
    ent = np.zeros((32, nmax8), dtype=np.int64)
    for lane, x in enumerate(lists):
        ent[lane, :len(x)] = x
    test = 0
    for k in range(nmax8):
        test += wavefronts(layout["h_base"] + layout["h_stride"] * ent[:, k], 8)
    stacks = stacks_override if stacks_override is not None else [x[m] for x, m in zip(lists, inside)]
    depth = max(len(x) for x in stacks)
    depth4 = (depth + 3) // 4 * 4
    drain = 0
    for r in range(depth4):
        j = np.array([x[len(x) - 1 - r] if r < len(x) else 0 for x in stacks] + [0] * (32 - len(stacks)))
        for base, stride, width in layout["drain"]:
            drain += wavefronts(base + stride * j, width)
    pushes = sum(len(x) for x in stacks)
    stack = int(np.ceil(pushes / 32.0 * 1.6)) + depth4            # pushes are predicated 2-byte stores (1 wavefront per pushing instruction), pops 1 per row
    return dict(test=test, drain=drain, stack=stack, entries=sum(len(x) for x in lists) / 32.0, pairs=pushes / 32.0, rows=depth4)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    pos, L = em.workloads.fcc_lattice(n)
    # a fluid, not a lattice: randomise the positions as the melted bench system would be (the lattice's symmetry biases bank patterns)
    rng = np.random.default_rng(3)
    pos = pos + rng.normal(scale=0.12, size=pos.shape)
    M, order, start = build_bricks(pos, L)
    cap = 2500
    layouts = {
        "in tree: xy double2[], z double[], fp16 uint2[]": dict(
            h_base=(cap + 1) * 24, h_stride=8, drain=[(0, 16, 16), ((cap + 1) * 16, 8, 8)]),
        "SoA x[], y[], z[] (3 x LDS.64)": dict(
            h_base=(cap + 1) * 24, h_stride=8, drain=[(0, 8, 8), ((cap + 1) * 8, 8, 8), ((cap + 1) * 16, 8, 8)]),
        "32-byte records {x,y | z,fp16}: 2 x LDS.128": dict(
            h_base=24, h_stride=32, drain=[(0, 32, 16), (16, 32, 16)]),
        "in tree + class-aware pops (16 sub-stacks per lane)": dict(
            h_base=(cap + 1) * 24, h_stride=8, drain=[(0, 16, 16), ((cap + 1) * 16, 8, 8)], drain_order=16),
        "in tree + list entries ordered by class at build time": dict(
            h_base=(cap + 1) * 24, h_stride=8, drain=[(0, 16, 16), ((cap + 1) * 16, 8, 8)], list_order=(16, False)),
        "in tree + ordered list (padded with dummies)": dict(
            h_base=(cap + 1) * 24, h_stride=8, drain=[(0, 16, 16), ((cap + 1) * 16, 8, 8)], list_order=(16, True)),
        "ordered list + class-aware pops": dict(
            h_base=(cap + 1) * 24, h_stride=8, drain=[(0, 16, 16), ((cap + 1) * 16, 8, 8)], list_order=(16, False), drain_order=16),
    }
    tot = {k: dict(test=0, drain=0, stack=0, entries=0.0, pairs=0.0, rows=0, tasks=0) for k in layouts}
    for b in range(nb):
        hx0, hy0, hz0 = (int(rng.integers(0, M // BX)) * BX, int(rng.integers(0, M // BY)) * BY, int(rng.integers(0, M // BZ)) * BZ)
        nst, home, lists, inside = brick_streams(pos, L, M, order, start, hx0, hy0, hz0)
        for g0 in range(0, len(home), 32):
            ls, ins = lists[g0:g0 + 32], inside[g0:g0 + 32]
            while len(ls) < 32:
                ls, ins = ls + [np.zeros(0, dtype=np.int64)], ins + [np.zeros(0, dtype=bool)]
            for name, lay in layouts.items():
                c = task_cost(ls, ins, lay)
                for k, vv in c.items():
                    tot[name][k] += vv
                tot[name]["tasks"] += 1
    for name, t in tot.items():
        k = t["tasks"]
        print("%-52s per task: test %6.0f  drain %6.0f  stack %5.0f  total %6.0f   (entries/atom %.1f, survivors/atom %.1f, drain rows %.1f; "
              "%.2f wavefronts per test LDS, %.2f per drain row)" % (name, t["test"] / k, t["drain"] / k, t["stack"] / k,
              (t["test"] + t["drain"] + t["stack"]) / k, t["entries"] / k, t["pairs"] / k, t["rows"] / k,
              t["test"] / k / (np.ceil(t["entries"] / k / 8) * 8), t["drain"] / max(t["rows"], 1)))


if __name__ == "__main__":
    main()
