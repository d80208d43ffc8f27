# EmDee.jl shim over libemdee_b200.so -- keeps the reference's exported names on the nonbonded path
# and forwards each to one `ccall` of the C ABI in include/emdee_b200.h.
#
# WRITTEN BLIND: Julia is not installed in the build container nor on the GPU boxes (SURVEY F11), so
# this file has never been executed; it is deliberately thin and mechanical.  tests/test_abi.py
# checks that every symbol it names is declared by the header and exported by the library.
# The runnable host mirror with identical semantics is emdee.jl_b200/api.py (Python + ctypes).
#
# What changes for a caller of the reference (see INTEGRATION.md):
#   * arrays are host `Array{Float64}` (no CUDA.jl in the loop); the library owns the device mirrors;
#   * `compute_nonbonded!` takes the same nine positional arguments; `tiles` is a host
#     Vector{Tuple{Int32,Int32}}; a keyword `mode` selects the pair-set semantics (default: the
#     reference's own all-pairs behaviour);
#   * `Cells` has no `action_cells`/`reaction_cells` tables; `head`/`next` are not materialised;
#   * additive: `NonbondedSystem`, `step!` (velocity-Verlet), `pair_set`, `pair_set_digest`.
module EmDee

export LennardJonesModel, LennardJonesAtom,
       FORCES, ENERGIES, VIRIALS,
       nonbonded_computation_tiles, compute_nonbonded!, naively_compute_nonbonded!,
       Cells, update_cells!,
       NonbondedSystem, step!, pair_set_digest

const libemdee = get(ENV, "EMDEE_B200_LIB", "libemdee_b200.so")   # same idiom as src/molecular_graphs.jl:4

const FORCES = 1 << 0          # src/nonbonded.jl:12-14
const ENERGIES = 1 << 1
const VIRIALS = 1 << 2
const WARPSIZE = 32            # src/nonbonded.jl:16
const CUTOFF = Cint(0)
const ALLPAIRS_REFERENCE = Cint(1)

function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:emdee_last_error, libemdee), Cstring, ()))
    error("emdee_b200 status $status: $msg")
end

# ---- src/lennard_jones.jl:6-18 -----------------------------------------------------------------
struct LennardJonesModel
    cutoff::Float64
    switch::Float64
    rc²::Float64
    rs²::Float64
    δ⁻²::Float64
    LennardJonesModel(cutoff, switch) = new(cutoff, switch, cutoff^2, switch^2, 1/(cutoff^2 - switch^2))
end

struct LJAtom
    half_σ::Float64
    twice_sqrt_ε::Float64
end
LennardJonesAtom(ε, σ) = LJAtom(0.5σ, 2*sqrt(ε))

# ---- src/nonbonded.jl:18-26 (integer work, identical) ---------------------------------------------
function nonbonded_computation_tiles(N)
    n = Int32(cld(N, WARPSIZE))
    pairs = Vector{Tuple{Int32,Int32}}(undef, n*(n+1)÷2)
    k = 0
    for i = 0:n-1, j = 1:n-i
        pairs[k+=1] = (j, j+i)
    end
    return pairs
end

# ---- handles ---------------------------------------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:emdee_create, libemdee), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device))
        ctx = new(h[])
        finalizer(c -> ccall((:emdee_destroy, libemdee), Cint, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end

const default_context = Ref{Union{Nothing,Context}}(nothing)
context() = (default_context[] === nothing && (default_context[] = Context()); default_context[])

mutable struct NonbondedSystem
    handle::Ptr{Cvoid}
    N::Int
    L::Float64
    ctx::Context
    function NonbondedSystem(N::Integer, L::Real; ctx::Context=context())
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:emdee_system_create, libemdee), Cint, (Ptr{Cvoid}, Int64, Cdouble, Ref{Ptr{Cvoid}}),
                    ctx.handle, N, L, h))
        s = new(h[], N, L, ctx)
        finalizer(x -> ccall((:emdee_system_destroy, libemdee), Cint, (Ptr{Cvoid},), x.handle), s)
        return s
    end
end

set_model!(s::NonbondedSystem, m::LennardJonesModel) =
    check(ccall((:emdee_set_model, libemdee), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), s.handle, m.cutoff, m.switch))
set_atoms!(s::NonbondedSystem, atoms::Vector{LJAtom}) =      # Vector{LJAtom} is 2xN Float64 in memory
    check(ccall((:emdee_set_lj_atoms, libemdee), Cint, (Ptr{Cvoid}, Ptr{LJAtom}), s.handle, atoms))
set_positions!(s::NonbondedSystem, r::Matrix{Float64}) =     # 3xN column-major, src/nonbonded.jl:60
    check(ccall((:emdee_set_positions, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, r))
set_velocities!(s::NonbondedSystem, v::Matrix{Float64}) =
    check(ccall((:emdee_set_velocities, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, v))
set_masses!(s::NonbondedSystem, m::Vector{Float64}) =
    check(ccall((:emdee_set_masses, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, m))
set_exclusions!(s::NonbondedSystem, base::Vector{Int32}, mask::Vector{UInt64}) =
    check(ccall((:emdee_set_exclusions, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{UInt64}), s.handle, base, mask))
# lj14scale of the force field (src/modelling.jl:199): pairs three bonds apart, 0-based ids with i < j, 2 x n column-major
set_pairs14!(s::NonbondedSystem, ij::Matrix{Int32}, scale::Real) =
    check(ccall((:emdee_set_pairs14, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}, Int64, Cdouble), s.handle, ij, size(ij, 2), scale))
set_skin!(s::NonbondedSystem, skin::Real) =
    check(ccall((:emdee_set_skin, libemdee), Cint, (Ptr{Cvoid}, Cdouble), s.handle, skin))
set_tiles!(s::NonbondedSystem, tiles::Vector{Tuple{Int32,Int32}}) =
    check(ccall((:emdee_set_tiles, libemdee), Cint, (Ptr{Cvoid}, Ptr{Tuple{Int32,Int32}}, Int64), s.handle, tiles, length(tiles)))
bin!(s::NonbondedSystem, ndiv::Integer=2) =
    check(ccall((:emdee_bin, libemdee), Cint, (Ptr{Cvoid}, Cint), s.handle, ndiv))
compute!(s::NonbondedSystem, mode::Integer, bitmask::Integer) =
    check(ccall((:emdee_compute_nonbonded, libemdee), Cint, (Ptr{Cvoid}, Cint, Cint), s.handle, mode, bitmask))
# the reference's call shape on a device-resident system: evaluation + the selected outputs into host arrays, as one call
# (on one GPU the rows of a finished chunk of z planes travel to the host while the next chunk computes)
compute_into!(forces::Matrix{Float64}, energies::Vector{Float64}, virials::Vector{Float64}, s::NonbondedSystem,
              mode::Integer, bitmask::Integer) =
    check(ccall((:emdee_compute_nonbonded_into, libemdee), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                s.handle, mode, bitmask, forces, energies, virials))
get_forces!(out::Matrix{Float64}, s::NonbondedSystem) =
    check(ccall((:emdee_get_forces, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
get_energies!(out::Vector{Float64}, s::NonbondedSystem) =
    check(ccall((:emdee_get_energies, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
get_virials!(out::Vector{Float64}, s::NonbondedSystem) =
    check(ccall((:emdee_get_virials, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
get_positions!(out::Matrix{Float64}, s::NonbondedSystem) =
    check(ccall((:emdee_get_positions, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
get_velocities!(out::Matrix{Float64}, s::NonbondedSystem) =
    check(ccall((:emdee_get_velocities, libemdee), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))

function totals(s::NonbondedSystem)
    E = Ref{Cdouble}(0); W = Ref{Cdouble}(0); n = Ref{Int64}(0)
    check(ccall((:emdee_get_totals, libemdee), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Int64}), s.handle, E, W, n))
    return E[], W[], n[]
end

function pair_set_digest(s::NonbondedSystem)
    d = zeros(UInt64, 3)
    check(ccall((:emdee_pair_set_digest, libemdee), Cint, (Ptr{Cvoid}, Ptr{UInt64}), s.handle, d))
    return d
end

# pairs (i<j) the pair-list stepping kernel evaluates inside the cutoff at the current positions (-1: no list in use)
function list_pair_count(s::NonbondedSystem)
    n = Ref{Int64}(0)
    check(ccall((:emdee_list_pair_count, libemdee), Cint, (Ptr{Cvoid}, Ref{Int64}), s.handle, n))
    return n[]
end

# velocity-Verlet (additive; the reference has no integrator).  rebin_every > 0: re-bin at that cadence; 0: never;
# < 0 (with a skin, set_skin!): adaptively, on the step on which an atom has moved more than skin/2 since the last binning
step!(s::NonbondedSystem, nsteps::Integer; dt::Real=0.005, rebin_every::Integer=1) =
    check(ccall((:emdee_vv_step, libemdee), Cint, (Ptr{Cvoid}, Cdouble, Int64, Cint), s.handle, dt, nsteps, rebin_every))

# thermostat pieces (additive): kinetic energy from the device, velocity rescaling on the device
function kinetic_energy(s::NonbondedSystem)
    K = Ref{Cdouble}(0)
    check(ccall((:emdee_kinetic_energy, libemdee), Cint, (Ptr{Cvoid}, Ref{Cdouble}), s.handle, K))
    return K[]
end
scale_velocities!(s::NonbondedSystem, factor::Real) =
    check(ccall((:emdee_scale_velocities, libemdee), Cint, (Ptr{Cvoid}, Cdouble), s.handle, factor))
# one Berendsen rescaling towards kT with coupling time tau after `elapsed` time of dynamics
function berendsen!(s::NonbondedSystem, kT::Real, tau::Real, elapsed::Real; ndof::Integer=3*s.N-3)
    now = 2*kinetic_energy(s)/ndof
    lambda = now > 0 ? sqrt(max(0.0, 1 + elapsed/tau*(kT/now - 1))) : 1.0
    scale_velocities!(s, lambda)
    return now, lambda
end
# how the stepping path is configured after bin!: (brick cells x,y,z, brick capacity, pair list?, persistent?, fused VV?, list chunks)
synchronize(s::NonbondedSystem) = check(ccall((:emdee_synchronize, libemdee), Cint, (Ptr{Cvoid},), s.handle))
# ids sorted by (cell, id) and the first slot of every cell (0-based), the order `Cells` is built from
function cell_order(s::NonbondedSystem, M::Integer)
    perm = zeros(Int32, s.N); start = zeros(Int32, M^3 + 1)
    check(ccall((:emdee_get_cell_order, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), s.handle, perm, start))
    return perm, start
end
# re-binnings and stepping launches by list mode (full walk, prune, replay) since the system was created
function step_counters(s::NonbondedSystem)
    o = zeros(Int64, 4)
    check(ccall((:emdee_get_step_counters, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int64}), s.handle, o))
    return o
end

function step_config(s::NonbondedSystem)
    o = zeros(Int32, 8)
    check(ccall((:emdee_get_step_config, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}), s.handle, o))
    return o
end

# ---- compute_nonbonded!, src/nonbonded.jl:109-120 ----------------------------------------------------
function compute_nonbonded!(forces::Matrix{Float64}, energies::Vector{Float64}, virials::Vector{Float64},
                            positions::Matrix{Float64}, L, tiles, model::LennardJonesModel,
                            atoms::Vector{LJAtom}, ::Val{bitmask};
                            mode::Integer=ALLPAIRS_REFERENCE, ndiv::Integer=2) where {bitmask}
    N = size(positions, 2)
    check(ccall((:emdee_compute_nonbonded_host, libemdee), Cint,
                (Int64, Ptr{Cdouble}, Cdouble, Cdouble, Cdouble, Ptr{LJAtom}, Ptr{Tuple{Int32,Int32}}, Int64,
                 Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                N, positions, Float64(L), model.cutoff, model.switch, atoms,
                tiles === nothing ? C_NULL : tiles, tiles === nothing ? 0 : length(tiles),
                mode, ndiv, bitmask, forces, energies, virials))
    return nothing
end

# naively_compute_nonbonded!, src/nonbonded.jl:122-155: all pairs, all three outputs, no tiles.
naively_compute_nonbonded!(forces, energies, virials, positions, L, model, atoms) =
    compute_nonbonded!(forces, energies, virials, positions, L, nothing, model, atoms,
                       Val(FORCES | ENERGIES | VIRIALS))

# ---- Cells / update_cells!, src/cells.jl:6-20,176-222 --------------------------------------------------
mutable struct Cells
    M::Int32
    cutoff::Float64
    ndiv::Int
    index::Vector{Int32}          # 1-based cell of every atom, src/cells.jl:85
    population::Vector{Int32}
    system::NonbondedSystem
end

function refresh!(cells::Cells, r::Matrix{Float64})
    s = cells.system
    set_positions!(s, r)
    bin!(s, cells.ndiv)
    M = Ref{Int32}(0)
    check(ccall((:emdee_get_cells_per_dimension, libemdee), Cint, (Ptr{Cvoid}, Ref{Int32}), s.handle, M))
    cells.M = M[]
    cells.index = Vector{Int32}(undef, s.N)
    check(ccall((:emdee_get_cell_index, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}), s.handle, cells.index))
    cells.population = Vector{Int32}(undef, Int(M[])^3)
    check(ccall((:emdee_get_cell_population, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}), s.handle, cells.population))
    return cells
end

function Cells(r::Matrix{Float64}, L, cutoff; ndiv=2, num_threads=256)
    s = NonbondedSystem(size(r, 2), L)
    set_model!(s, LennardJonesModel(cutoff, cutoff/2))
    return refresh!(Cells(0, cutoff, ndiv, Int32[], Int32[], s), r)
end

# update_cells!(cells, r, L), src/cells.jl:196-222: incremental -- the device counts the atoms whose cell changed and re-sorts
# only if there are any; the host arrays are refreshed in that case only
function update_cells!(cells::Cells, r::Matrix{Float64}, L)
    s = cells.system
    set_positions!(s, r)
    movers = Ref{Int64}(0)
    check(ccall((:emdee_update_cells, libemdee), Cint, (Ptr{Cvoid}, Ref{Int64}), s.handle, movers))
    if movers[] != 0
        check(ccall((:emdee_get_cell_index, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}), s.handle, cells.index))
        check(ccall((:emdee_get_cell_population, libemdee), Cint, (Ptr{Cvoid}, Ptr{Int32}), s.handle, cells.population))
    end
    return nothing
end

# Slab ranks (multi-GPU): move only the rows of the full id-ordered host arrays that belong to atoms this rank owns -- a cyclic
# window first, first+1, ... (mod N) of `count` rows; the arrays passed are the FULL 3xN / N arrays
function local_id_range(s::NonbondedSystem)
    a = Ref{Int64}(0); n = Ref{Int64}(0)
    check(ccall((:emdee_get_local_id_range, libemdee), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), s.handle, a, n))
    return a[], n[]
end
set_positions_range!(s::NonbondedSystem, id_first::Integer, count::Integer, r::Matrix{Float64}) =
    check(ccall((:emdee_set_positions_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, count, r))
forces_range!(out::Matrix{Float64}, s::NonbondedSystem, id_first::Integer, count::Integer) =
    check(ccall((:emdee_get_forces_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, count, out))
energies_range!(out::Vector{Float64}, s::NonbondedSystem, id_first::Integer, count::Integer) =
    check(ccall((:emdee_get_energies_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, count, out))
virials_range!(out::Vector{Float64}, s::NonbondedSystem, id_first::Integer, count::Integer) =
    check(ccall((:emdee_get_virials_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, count, out))

end
set_positions_range!(s::NonbondedSystem, id_first::Integer, rows::Matrix{Float64}) =
    check(ccall((:emdee_set_positions_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, size(rows, 2), rows))
forces_range!(out::Matrix{Float64}, s::NonbondedSystem, id_first::Integer) =
    check(ccall((:emdee_get_forces_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, size(out, 2), out))
energies_range!(out::Vector{Float64}, s::NonbondedSystem, id_first::Integer) =
    check(ccall((:emdee_get_energies_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, length(out), out))
virials_range!(out::Vector{Float64}, s::NonbondedSystem, id_first::Integer) =
    check(ccall((:emdee_get_virials_range, libemdee), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), s.handle, id_first, length(out), out))

end
