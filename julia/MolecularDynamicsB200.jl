# MolecularDynamicsB200.jl -- the reference-side binding of include/mdb200.h.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia (SURVEY.md F4).  It is the `ccall` layer a
# maintainer of MolecularDynamics.jl adds so that the existing API
#     Parameters(...) -> initialize_state(...) -> state.velocities = initialize_velocities(...) -> run_simulation!(...)
# keeps working with the per-step loop (src/simulation.jl:88-108, :231-250) executed by libmdb200.so on a B200.
# INTEGRATION.md walks through it.  The Python package moleculardynamics.jl_b200/ mirrors this file one to one and is
# what tests/ and bench.py drive.
module MolecularDynamicsB200

using MolecularDynamics
using MolecularDynamics: Parameters, SimulationState, Ensemble, NVE, NVT, Brownian, Potential, PseudoHS, LennardJones,
                         LennardJonesXPLOR, energy_lrc, pressure_lrc, compute_box_volume, open_files,
                         write_to_file_lammps, finalize_simulation!, generate_log_times
using StaticArrays, Printf, LinearAlgebra

const libmdb = get(ENV, "MDB200_LIB", "libmdb200.so")

# ---- include/mdb200.h mirrored -------------------------------------------------------------------------------
struct MdbConfig
    dim::Int32
    potential::Int32
    n_particles::Int64
    unitcell::NTuple{9,Float64}
    cutoff::Float64
    pot_params::NTuple{8,Float64}
    seed::UInt64
    device::Int32
    mode::Int32
    skin::Float64
    use_graph::Int32
    rank::Int32
    nranks::Int32
    no_fuse::Int32
    skin_inner::Float64
    slab_transport::Int32      # nranks > 1: 0 default (peer memory, else NCCL), 1 classic NCCL send/recv, 2 peer memory
    reserved::Int32
end

const MDB_OK = Cint(0)
const Handle = Ptr{Cvoid}

struct MdbError <: Exception
    code::Cint
    msg::String
end

function check(h::Handle, rc::Cint)
    rc == MDB_OK && return nothing
    msg = unsafe_string(ccall((:mdb_last_error, libmdb), Cstring, (Handle,), h))
    throw(MdbError(rc, msg))
end

# Potential subtype -> (device functor tag, parameters): the plugin contract of src/types.jl:1-6.
# A subtype without a method here has no device functor; like the reference's fallback `evaluate` it is an error.
potential_tag(::PseudoHS) = (Int32(0), ())
potential_tag(p::LennardJones) = (Int32(1), (p.epsilon, p.r_cut))
potential_tag(p::LennardJonesXPLOR) = (Int32(2), (p.ϵ, p.r_on, p.r_cut))
potential_tag(p::Potential) = error("evaluate not implemented on the device for potential type: $(typeof(p))")
# The README's user-defined plugin (README.md:82-145) is a struct the USER defines, so its tag cannot be a method on a type of
# this module; the user registers it with one line next to the struct (device functor 3 = PotPoly, params {rcut, non_additivity}):
#   MolecularDynamicsB200.potential_tag(p::Polydisperse) = MolecularDynamicsB200.polydisperse_tag(p.rcut, p.non_additivity)
polydisperse_tag(rcut::Real=1.25, non_additivity::Real=0.2) = (Int32(3), (Float64(rcut), Float64(non_additivity)))

pad8(t) = ntuple(i -> i <= length(t) ? Float64(t[i]) : 0.0, 8)

"""
    GPUSystem

Takes the place of `CellListMap.ParticleSystem` in `SimulationState.system` (src/types.jl:15-17).  Exposes the
property names the drivers touch (`positions`, `xpositions`, `energy_and_forces`) but the data lives in HBM.
"""
mutable struct GPUSystem{D}
    handle::Handle
    n::Int
    cutoff::Float64
    function GPUSystem{D}(cfg::MdbConfig) where {D}
        h = Ref{Handle}(C_NULL)
        rc = ccall((:mdb_create, libmdb), Cint, (Ref{MdbConfig}, Ref{Handle}), cfg, h)
        rc == MDB_OK || throw(MdbError(rc, unsafe_string(ccall((:mdb_last_error, libmdb), Cstring, (Handle,), C_NULL))))
        sys = new{D}(h[], Int(cfg.n_particles), cfg.cutoff)
        finalizer(s -> ccall((:mdb_destroy, libmdb), Cint, (Handle,), s.handle), sys)
        return sys
    end
end

# Vector{MVector{D,Float64}} <-> the AoS host image the C ABI takes (zero-copy for SVector, one copy for MVector)
flat(v::Vector{<:StaticVector{D,T}}) where {D,T} = T[x[k] for x in v for k in 1:D]
unflat(::Val{D}, a::Vector{T}) where {D,T} = [MVector{D,T}(ntuple(k -> a[(i - 1) * D + k], D)) for i in 1:(length(a) ÷ D)]

function upload!(sys::GPUSystem{D}, positions, diameters; velocities=nothing, forces=nothing, images=nothing) where {D}
    p(x) = x === nothing ? C_NULL : pointer(x)
    x, d = flat(positions), Vector{Float64}(diameters)
    v = velocities === nothing ? nothing : flat(velocities)
    f = forces === nothing ? nothing : flat(forces)
    im = images === nothing ? nothing : flat(images)
    GC.@preserve x d v f im check(sys.handle, ccall((:mdb_upload, libmdb), Cint,
        (Handle, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}), sys.handle, p(x), p(v), p(f), p(d), p(im)))
end

function download(sys::GPUSystem{D}; positions=true, velocities=true, forces=true, images=true) where {D}
    x = positions ? Vector{Float64}(undef, D * sys.n) : nothing
    v = velocities ? Vector{Float64}(undef, D * sys.n) : nothing
    f = forces ? Vector{Float64}(undef, D * sys.n) : nothing
    im = images ? Vector{Int32}(undef, D * sys.n) : nothing
    p(a) = a === nothing ? C_NULL : pointer(a)
    GC.@preserve x v f im check(sys.handle, ccall((:mdb_download, libmdb), Cint,
        (Handle, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}), sys.handle, p(x), p(v), p(f), p(im)))
    u(a) = a === nothing ? nothing : unflat(Val(D), a)
    return u(x), u(v), u(f), u(im)
end

# Slab handles (cfg.nranks > 1, one Julia process per GPU): only the rows a rank owns travel between host and device once
# mdb_upload has planned the slabs.  `ids` are the original (0-based) particle indices of the rows.
function download_owned(sys::GPUSystem{D}; capacity::Int=sys.n) where {D}
    ids = Vector{Int32}(undef, capacity)
    x, v, f = (Vector{Float64}(undef, D * capacity) for _ in 1:3)
    im = Vector{Int32}(undef, D * capacity)
    cnt = Ref{Int64}(0)
    GC.@preserve ids x v f im check(sys.handle, ccall((:mdb_download_owned, libmdb), Cint,
        (Handle, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int64}),
        sys.handle, capacity, ids, x, v, f, im, cnt))
    k = Int(cnt[])
    return ids[1:k], unflat(Val(D), x[1:D*k]), unflat(Val(D), v[1:D*k]), unflat(Val(D), f[1:D*k]), unflat(Val(D), im[1:D*k])
end
function upload_owned!(sys::GPUSystem{D}, ids::Vector{Int32}, positions, diameters; velocities=nothing, forces=nothing, images=nothing) where {D}
    p(a) = a === nothing ? C_NULL : pointer(a)
    x, d = flat(positions), Vector{Float64}(diameters)
    v = velocities === nothing ? nothing : flat(velocities)
    f = forces === nothing ? nothing : flat(forces)
    im = images === nothing ? nothing : flat(images)
    GC.@preserve ids x d v f im check(sys.handle, ccall((:mdb_upload_owned, libmdb), Cint,
        (Handle, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        sys.handle, length(ids), ids, p(x), p(v), p(f), p(d), p(im)))
end

struct EnergyAndForcesView{D}
    sys::GPUSystem{D}
end
function thermo(sys::GPUSystem)
    out = zeros(4)
    check(sys.handle, ccall((:mdb_thermo, libmdb), Cint, (Handle, Ptr{Float64}), sys.handle, out))
    return out
end
function Base.getproperty(sys::GPUSystem{D}, name::Symbol) where {D}
    name === :positions || name === :xpositions ? download(sys; velocities=false, forces=false, images=false)[1] :
    name === :energy_and_forces ? EnergyAndForcesView{D}(sys) : getfield(sys, name)
end
function Base.getproperty(v::EnergyAndForcesView, name::Symbol)
    sys = getfield(v, :sys)
    name === :forces ? download(sys; positions=false, velocities=false, images=false)[3] :
    name === :energy ? thermo(sys)[1] : name === :virial ? thermo(sys)[2] : getfield(v, name)
end

"map_pairwise!(energy_and_forces!, system) of src/simulation.jl:99-104 on the device"
function map_pairwise!(sys::GPUSystem)
    e, w, n = Ref(0.0), Ref(0.0), Ref(Int64(0))
    check(sys.handle, ccall((:mdb_compute_forces, libmdb), Cint, (Handle, Ref{Float64}, Ref{Float64}, Ref{Int64}), sys.handle, e, w, n))
    return e[], w[], n[]
end

"""
    to_gpu(state, params; cutoff=1.5, seed=rand(UInt64), device=0, mode=0, skin=0.0)

Move a `SimulationState` produced by the stock `initialize_state` (src/initialization.jl:112-157) to the GPU: the
`system` field is replaced by a `GPUSystem` holding positions, velocities (if already assigned), forces, images.
"""
function to_gpu(state::SimulationState, params::Parameters; cutoff=1.5, seed=rand(UInt64), device=0, mode=0, skin=0.0)
    D = state.dimension
    U = state.unitcell   # diagonal or general (to_unitcell's matrix branch, src/initialization.jl:13-15): the engine wraps,
                         # images and bins through the fractional coordinates for a cell with off-diagonal entries
    tag, pp = potential_tag(params.potential)
    cell = ntuple(q -> (r = (q - 1) ÷ 3 + 1; c = (q - 1) % 3 + 1; (r <= D && c <= D) ? Float64(U[r, c]) : 0.0), 9)
    cfg = MdbConfig(D, tag, length(state.system.xpositions), cell, cutoff, pad8(pp), seed, device, mode, skin, 1, 0, 1,
                    Int32(0), 0.0, Int32(0), Int32(0))
    sys = GPUSystem{D}(cfg)
    upload!(sys, state.system.xpositions, state.diameters;
            velocities=isempty(state.velocities) ? nothing : state.velocities,
            forces=state.system.energy_and_forces.forces, images=state.images)
    return SimulationState(sys, state.diameters, state.rng, state.unitcell, state.velocities, state.images, D, state.nf)
end

set_velocities!(sys::GPUSystem, v) = (a = flat(v); GC.@preserve a check(sys.handle,
    ccall((:mdb_set_velocities, libmdb), Cint, (Handle, Ptr{Float64}), sys.handle, a)))

function run_chunk!(sys::GPUSystem, ::NVE, params, steps, first_step)
    t = zeros(4, length(steps))
    check(sys.handle, ccall((:mdb_run_nve, libmdb), Cint, (Handle, Int64, Float64, Ptr{Float64}), sys.handle, length(steps), params.dt, t))
    return t
end
function run_chunk!(sys::GPUSystem, ens::NVT, params, steps, first_step)
    kt = Float64[ens.ktemp(s + 1) for s in steps]          # ensemble.ktemp(step + 1), src/simulation.jl:108, src/integrate.jl:49
    t = zeros(4, length(steps))
    check(sys.handle, ccall((:mdb_run_nvt, libmdb), Cint, (Handle, Int64, Float64, Ptr{Float64}, Float64, Ptr{Float64}),
        sys.handle, length(steps), params.dt, kt, ens.tau, t))
    return t
end
function run_chunk!(sys::GPUSystem, ens::Brownian, params, steps, first_step)
    t = zeros(4, length(steps))
    check(sys.handle, ccall((:mdb_run_brownian, libmdb), Cint, (Handle, Int64, Float64, Float64, Ptr{Float64}),
        sys.handle, length(steps), params.dt, ens.ktemp, t))
    return t
end

# initialize_velocities (src/initialization.jl:32-47) on the device; `stream` selects an independent set of draws
init_velocities!(sys::GPUSystem, ktemp::Float64; stream::Integer=0) =
    check(sys.handle, ccall((:mdb_init_velocities, libmdb), Cint, (Handle, Float64, UInt64), sys.handle, ktemp, UInt64(stream)))
# mdb_fire_params (include/mdb200.h) and the FIRE minimiser, second caller of the force path (src/minimize.jl:31-135)
struct FireParams
    max_steps::Int64
    tol::Float64
    dt_initial::Float64
    dt_max::Float64
    alpha0::Float64
    f_inc::Float64
    f_dec::Float64
    n_min::Int32
    reserved::Int32
end
"returns (energy, F_rms, steps_done, converged) like fire_minimize! reports them (src/minimize.jl:127-135)"
function fire_minimize!(sys::GPUSystem; max_steps::Int=10_000, tol::Float64=1e-6, dt_initial::Float64=0.01, dt_max::Float64=0.1,
                        alpha0::Float64=0.1, f_inc::Float64=1.2, f_dec::Float64=0.2, n_min::Int=5)
    fp = Ref(FireParams(max_steps, tol, dt_initial, dt_max, alpha0, f_inc, f_dec, Int32(n_min), Int32(0)))
    out, conv = zeros(3), Ref{Int32}(0)
    check(sys.handle, ccall((:mdb_fire_minimize, libmdb), Cint, (Handle, Ptr{FireParams}, Ptr{Float64}, Ptr{Int32}), sys.handle, fp, out, conv))
    return out[1], out[2], Int(out[3]), conv[] != 0
end

# initialize_random (src/initialization.jl:20-30) on the GPU: uniform points of the cell, then FIRE on the penalty
# potential MDB_POT_SOFT until no pair is closer than `tol` (what Packmol.pack_monoatomic! does on the host)
function initialize_random_gpu(unitcell, npart::Int, dimension::Int; tol::Float64=1.0, seed::UInt64=rand(UInt64), device::Int=0)
    cell = ntuple(q -> (r = (q - 1) ÷ 3 + 1; c = (q - 1) % 3 + 1; (r <= dimension && c <= dimension) ? Float64(unitcell[r, c]) : 0.0), 9)
    tp = 1.001 * tol
    cfg = MdbConfig(dimension, Int32(4), npart, cell, tp, pad8((1.0, tp)), seed, device, 0, 0.0, 1, 0, 1, Int32(0), 0.0,
                    Int32(0), Int32(0))
    sys = dimension == 3 ? GPUSystem{3}(cfg) : GPUSystem{2}(cfg)
    upload!(sys, [zeros(dimension) for _ in 1:npart], ones(npart); velocities=[zeros(dimension) for _ in 1:npart])
    check(sys.handle, ccall((:mdb_random_positions, libmdb), Cint, (Handle, UInt64), sys.handle, UInt64(0)))
    fp = Ref(FireParams(2000, 1e-12, 0.02, 0.2, 0.1, 1.2, 0.2, Int32(5), Int32(0)))
    out, conv, close = zeros(3), Ref{Int32}(0), Ref{Int64}(1)
    while close[] != 0
        check(sys.handle, ccall((:mdb_fire_minimize, libmdb), Cint, (Handle, Ptr{FireParams}, Ptr{Float64}, Ptr{Int32}), sys.handle, fp, out, conv))
        check(sys.handle, ccall((:mdb_count_pairs, libmdb), Cint, (Handle, Float64, Ptr{Int64}, Ptr{Int32}), sys.handle, tol, close, C_NULL))
    end
    x, _, _, _ = download(sys; velocities=false, forces=false, images=false)
    return x
end
# exact binary restart: the saved run and a run restored from the file continue bit-identically
save_checkpoint(sys::GPUSystem, path::String) =
    check(sys.handle, ccall((:mdb_checkpoint_save, libmdb), Cint, (Handle, Cstring), sys.handle, path))
load_checkpoint!(sys::GPUSystem, path::String) =
    check(sys.handle, ccall((:mdb_checkpoint_load, libmdb), Cint, (Handle, Cstring), sys.handle, path))

"""
    run_simulation!(state::SimulationState{<:GPUSystem}, params, ensemble, total_steps, frequency, pathname; ...)

Same signature, files and thermo rows as src/simulation.jl:40-178 / :181-308.  The loop body runs on the GPU in chunks that
end exactly at the steps where the reference writes output; only then is state brought back for the text writers.
"""
function MolecularDynamics.run_simulation!(state::SimulationState{<:GPUSystem}, params::Parameters, ensemble::Ensemble,
        total_steps::Int, frequency::Int, pathname::String; traj_name::String="trajectory.xyz",
        thermo_name::String="thermo.txt", compress::Bool=false, log_times::Bool=false)
    (trajectory_file, thermo_file) = open_files(pathname, traj_name, thermo_name)
    open(io -> println(io, "# Step Energy Temperature Pressure"), thermo_file, "a")
    sys, D, N = state.system, state.dimension, params.n_particles
    volume = compute_box_volume(state.unitcell)
    isempty(state.velocities) || set_velocities!(sys, state.velocities)     # lazily assigned velocities (SURVEY Q12)
    virial, nprom, done = 0.0, 0, 0
    # log-spaced snapshots (src/simulation.jl:81-85, 153-171): generate_log_times() (src/io.jl:17-36) with step 0 in front;
    # the reference compares ONE pending entry per step, so a snapshot step is an output stop like a `frequency` step
    snapshot_times = log_times ? insert!(generate_log_times(), 1, 0) : Int[]
    snap_index = 1
    stops = sort(unique(vcat(collect(0:frequency:(total_steps - 1)), filter(s -> s < total_steps, snapshot_times))))
    frame_slot = Int32(0)
    function emit_frame(path::String, step::Int, append::Bool)
        # write_to_file_lammps (src/io.jl:78-170) without bringing the state back: the frame (unwrapped coordinates
        # included) is packed on the device, copied on the engine's copy stream and written by its background thread
        # while the next chunk of steps runs
        check(sys.handle, ccall((:mdb_frame_capture, libmdb), Cint, (Handle, Int32), sys.handle, frame_slot))
        check(sys.handle, ccall((:mdb_frame_write_lammps, libmdb), Cint, (Handle, Int32, Cstring, Int64, Int32),
            sys.handle, frame_slot, path, step, append ? 1 : 0))
        frame_slot = Int32(1) - frame_slot
    end
    for step in stops
        t = run_chunk!(sys, ensemble, params, done:step, done)               # steps done..step inclusive
        if ensemble isa Brownian
            for (k, s) in enumerate(done:step)
                if mod(s, 10) == 0
                    virial += t[2, k]; nprom += 1                            # src/simulation.jl:253-256
                end
            end
        end
        done = step + 1
        if mod(step, frequency) == 0
            U, W, KE = t[1, end], t[2, end], t[3, end]
            if ensemble isa Brownian
                # nprom == 0 gives 0.0/0 = NaN in the reference too (src/simulation.jl:252-262)
                row = (step, U / N, ensemble.ktemp, virial / (D * nprom * volume) + params.ρ * ensemble.ktemp)
                virial, nprom = 0.0, 0
            else
                T = 2.0 * KE / state.nf
                row = (step, (U + energy_lrc(params.potential, N, volume)) / N, T,
                       W / (D * volume) + params.ρ * T + pressure_lrc(params.potential, N, volume))
            end
            open(io -> Printf.format(io, Printf.Format("%d %.6f %.6f %.6f\n"), row...), thermo_file, "a")
            emit_frame(trajectory_file, step, true)
        end
        if log_times && snap_index <= length(snapshot_times) && snapshot_times[snap_index] == step
            emit_frame(joinpath(pathname, "snapshot.$(step)"), step, false)
            snap_index += 1
        end
    end
    done < total_steps && run_chunk!(sys, ensemble, params, done:(total_steps - 1), done)
    check(sys.handle, ccall((:mdb_frame_flush, libmdb), Cint, (Handle,), sys.handle))
    _, v, _, img = download(sys; positions=false, forces=false)
    ensemble isa Brownian || (state.velocities = v)
    state.images = img
    finalize_simulation!(trajectory_file, pathname, total_steps, state, params, compress)
    return nothing
end

end # module
