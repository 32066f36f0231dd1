# make_reference_golden.jl -- pins the oracle and the GPU path to the REFERENCE ITSELF the moment a Julia toolchain is at hand.
#
# The reference (edwinb-ai/MolecularDynamics.jl) ships no tests or golden vectors and Julia is absent from the build image
# (SURVEY.md F2/F4), so today parity is anchored on the C restatement in oracle/ ("parity unpinned" in DESIGN.md section 5).
# This script closes that gap without touching the reference: it runs the STOCK package on the committed C1 / C2 snapshots
# and writes what tests/test_reference_golden.py consumes.  Nothing of this repository's engine is loaded.
#
#   python tests/golden/make_reference_inputs.py          # snapshots -> raw little-endian files in tests/golden/ref_inputs/
#   julia --project=/path/to/MolecularDynamics.jl -t 1 julia/make_reference_golden.jl /path/to/this/repo
#   python tests/golden/ref_to_npz.py                      # raw outputs -> tests/golden/ref_c1.npz, ref_c2.npz (commit them)
#
# What is computed, with the reference's own calls:
#   forces   reset_output! + CellListMap.map_pairwise!(energy_and_forces!) exactly as src/simulation.jl:99-104 does, on a
#            system built like src/initialization.jl:100-107 -> E, W, per-particle forces;
#   pairs    a second map_pairwise! with a counting closure: pairs with d2 <= cutoff^2 (what the cell list visits) and pairs for
#            which evaluate() is non-zero (interacting) -- the two integers the GPU must reproduce bit-exactly;
#   nve      50 velocity-Verlet steps with the loop body of src/simulation.jl:88-108 (integrate_half!, forces,
#            integrate_second_half!) from the snapshot's velocities -> final positions, velocities, images and the per-step
#            (U, W, KE) rows.  NVT/Brownian are not exported: their random streams (Xoshiro, Distributions.Gamma) are not
#            reproducible on the device by design (north_star prescribes a counter-based generator).
# -t 1: one thread makes CellListMap's summation order deterministic; the GPU/oracle comparison uses 1e-12 relative anyway.
using MolecularDynamics
using CellListMap
using StaticArrays
using LinearAlgebra
using Printf

const MD = MolecularDynamics
root = length(ARGS) >= 1 ? ARGS[1] : dirname(@__DIR__)
indir = joinpath(root, "tests", "golden", "ref_inputs")
outdir = joinpath(root, "tests", "golden", "ref_outputs")
mkpath(outdir)

readf64(path, dims...) = (a = Array{Float64}(undef, dims...); read!(path, a); a)
writeraw(path, a) = open(io -> write(io, a), path, "w")

# the README's user-defined plugin (README.md:82-145), needed for C2
struct Polydisperse <: MD.Potential
    rcut::Float64
    non_additivity::Float64
end
function poly_potential(r, σ; rcut=1.25)
    c0, c2, c4 = -28.0 / rcut^12, 48.0 / rcut^14, -21.0 / rcut^16
    if r < rcut * σ
        sr = σ / r
        rs = r / σ
        u = sr^12 + c0 + c2 * rs^2 + c4 * rs^4
        f = 12.0 * σ^12 / r^13 - 2.0 * c2 * r / σ^2 - 4.0 * c4 * r^3 / σ^4
        return (u, f)
    end
    return (0.0, 0.0)
end
function MD.evaluate(p::Polydisperse, r::Float64, σ1::Float64, σ2::Float64)
    σ = 0.5 * (σ1 + σ2) * (1.0 - p.non_additivity * abs(σ1 - σ2))
    return poly_potential(r, σ; rcut=p.rcut)
end

function build_system(x::Matrix{Float64}, box::Vector{Float64}, cutoff, D)
    positions = [MVector{D,Float64}(x[:, i]) for i in 1:size(x, 2)]
    unitcell = MD.to_unitcell(box[1], D)              # cubic / square snapshots (src/initialization.jl:7-18)
    forces = similar(positions)
    forces .= zero.(positions)
    eaf = MD.EnergyAndForces(zero(cutoff), zero(cutoff), forces)
    system = CellListMap.ParticleSystem(; xpositions=positions, unitcell=unitcell, cutoff=cutoff, output=eaf,
                                        output_name=:energy_and_forces, parallel=false)
    return system, unitcell
end

function forces!(system, diameters, pot)
    MD.reset_output!(system.energy_and_forces)
    CellListMap.map_pairwise!((x, y, i, j, d2, out) -> MD.energy_and_forces!(x, y, i, j, d2, diameters, out, pot), system)
    return system.energy_and_forces
end

function count_pairs(x, box, cutoff, D, diameters, pot)
    positions = [SVector{D,Float64}(x[:, i]) for i in 1:size(x, 2)]
    sys = CellListMap.ParticleSystem(; xpositions=positions, unitcell=MD.to_unitcell(box[1], D), cutoff=cutoff, output=[0, 0],
                                     output_name=:counts, parallel=false)
    CellListMap.map_pairwise!(sys) do xi, yj, i, j, d2, c
        c[1] += 1
        (u, f) = MD.evaluate(pot, sqrt(d2), diameters[i], diameters[j])
        (u != 0.0 || f != 0.0) && (c[2] += 1)
        return c
    end
    return sys.counts
end

function export_case(name, D, pot, cutoff, dt)
    n = parse(Int, strip(read(joinpath(indir, "$(name)_n.txt"), String)))
    x = readf64(joinpath(indir, "$(name)_x.f64"), D, n)          # column-major D x n == the C-order (n, D) array of numpy
    v = readf64(joinpath(indir, "$(name)_v.f64"), D, n)
    diam = vec(readf64(joinpath(indir, "$(name)_diam.f64"), n))
    box = vec(readf64(joinpath(indir, "$(name)_box.f64"), D))
    system, unitcell = build_system(x, box, cutoff, D)
    eaf = forces!(system, diam, pot)
    F = reduce(hcat, [Vector(f) for f in eaf.forces])
    writeraw(joinpath(outdir, "$(name)_F.f64"), F)
    writeraw(joinpath(outdir, "$(name)_EW.f64"), [eaf.energy, eaf.virial])
    c = count_pairs(x, box, cutoff, D, diam, pot)
    writeraw(joinpath(outdir, "$(name)_counts.i64"), Int64[c[1], c[2]])
    # NVE loop body of src/simulation.jl:88-108, forces NOT primed before step 0 (zero forces, src/initialization.jl:97-98)
    system, unitcell = build_system(x, box, cutoff, D)
    unitcell_inv = inv(unitcell)
    velocities = [MVector{D,Float64}(v[:, i]) for i in 1:n]
    images = [zeros(MVector{D,Int32}) for _ in 1:n]
    nsteps = 50
    rows = zeros(3, nsteps)
    for s in 1:nsteps
        MD.integrate_half!(system.positions, images, velocities, system.energy_and_forces.forces, dt, unitcell, unitcell_inv)
        e = forces!(system, diam, pot)
        MD.integrate_second_half!(velocities, system.energy_and_forces.forces, dt)
        ke = 0.5 * sum(dot(w, w) for w in velocities)
        rows[:, s] .= (e.energy, e.virial, ke)
    end
    writeraw(joinpath(outdir, "$(name)_nve_x.f64"), reduce(hcat, [Vector(p) for p in system.positions]))
    writeraw(joinpath(outdir, "$(name)_nve_v.f64"), reduce(hcat, [Vector(w) for w in velocities]))
    writeraw(joinpath(outdir, "$(name)_nve_img.i32"), reduce(hcat, [Vector(m) for m in images]))
    writeraw(joinpath(outdir, "$(name)_nve_thermo.f64"), rows)
    @printf("%s: n=%d E=%.17g W=%.17g pairs(cutoff)=%d pairs(interacting)=%d\n", name, n, eaf.energy, eaf.virial, c[1], c[2])
end

export_case("c1", 3, MD.PseudoHS(), 1.5, 1e-3)
export_case("c2", 2, Polydisperse(1.25, 0.2), 1.5, 5e-3)
println("wrote ", outdir, "; now run: python tests/golden/ref_to_npz.py")
