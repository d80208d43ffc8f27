cd /root/repo
mkdir -p gpurun_out
EMDEE_DEBUG=1 timeout 120 python - <<'PY' 2>&1 | tail -12
import numpy as np, sys
sys.path.insert(0, '.')
import emdee_jl_b200 as em
from oracle import oracle_c as oc
oc.build()
for n in (24, 40):
    pos, L = em.workloads.fcc_lattice(n)
    N = pos.shape[0]
    atoms = em.workloads.lj_fluid_atoms(N)
    s = em.NonbondedSystem(N, L)
    s.set_model(em.LennardJonesModel(2.5, 2.0)); s.set_atoms(atoms); s.set_positions(pos)
    s.set_velocities(em.workloads.maxwell_velocities(N, 1.44)); s.set_masses(np.ones(N))
    s.set_skin(0.45); s.bin(1); s.compute(em.CUTOFF, 7)
    ref = oc.cutoff_cells(pos, L, 2.5, 2.0, atoms, ndiv=1, fast=True)
    f = s.forces(); E, W, npairs = s.totals()
    print("n", n, "single point (EW, TMA?)", s.step_config(), "ferr", np.abs(f - ref["forces"]).max() / np.sqrt((ref["forces"]**2).sum(1).mean()), "Eerr", abs(E - ref["E"]) / abs(ref["E"]))
    s.compute(em.CUTOFF, em.FORCES)
    s.vv_step(0.005, 12, -1); s.synchronize()
    ref = oc.cutoff_cells(s.positions(), L, 2.5, 2.0, atoms, ndiv=1, fast=True)
    print("   after 12 steps ferr", np.abs(s.forces() - ref["forces"]).max() / np.sqrt((ref["forces"]**2).sum(1).mean()), "pairs", s.list_pair_count(), ref["npairs"])
    s.close()
PY
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
B="timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-iters 3"
EMDEE_DEBUG=1 EMDEE_TMA=1 $B > gpurun_out/tma_on.json 2> gpurun_out/tma_on.err; grep "TMA staging" gpurun_out/tma_on.err | tail -1; tail -1 gpurun_out/tma_on.err | cut -c1-300
EMDEE_TMA=0 $B > gpurun_out/tma_off.json 2> gpurun_out/tma_off.err
for v in off on; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/tma_$v.json') if l.startswith('{')][-1]); print('TMA $v: ms/step %.4f kernel %.4f build %.4f parity %s frac %.4f e2e %.2f ms'%(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['list_build']['ms_per_launch'], d['parity']['ok'] if d['parity'] else None, d['roofline']['frac'], d['e2e']['ms_per_call']))"; done
