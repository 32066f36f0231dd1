"""Host-side mirror of the reference's public interface for the hot path.

Same names, argument meaning and error behaviour as /root/reference/src (Julia); the Julia shim with the same
shape is julia/MolecularDynamicsB200.jl.  Only what sits on or directly around the per-step path is mirrored:
Parameters / Potential subtypes / ensembles / initialize_state / initialize_velocities / run_simulation! and the
thermo + trajectory rows the loop writes.  All physics runs in csrc/libmdb200.so (CUDA); nothing here computes
forces or integrates on the CPU.
"""
import math
import os
import re

import numpy as np

from . import _capi


# ---- src/types.jl:1-6, src/potentials.jl -------------------------------------------------------------------
class Potential:
    """abstract type Potential (src/types.jl:1).  Subtypes carry `tag` + `params()` selecting the device functor."""
    tag = None

    def params(self):
        return ()

    def tail_correction_enabled(self):
        return False


def evaluate(pot, r, sigma1=None, sigma2=None, **kwargs):
    """Fallback of src/types.jl:4-6: a Potential without a device functor is an error (no CPU evaluation path)."""
    raise NotImplementedError("evaluate not implemented for potential type: %s" % type(pot).__name__)


class PseudoHS(Potential):
    """src/potentials.jl:5-29"""
    tag = _capi.POT_PSEUDOHS


class LennardJones(Potential):
    """src/potentials.jl:41-64 (keyword constructor); evaluate always uses lj_unshifted (:160-164)."""
    tag = _capi.POT_LJ

    def __init__(self, epsilon=1.0, sigma=1.0, r_cut=2.5, shift=False, force_shift=False, tail_correction=False):
        self.epsilon, self.sigma, self.r_cut = float(epsilon), float(sigma), float(r_cut)
        self.shift, self.force_shift, self.tail_correction = shift, force_shift, tail_correction
        srcut = sigma / r_cut
        srcut2 = srcut * srcut
        srcut6 = srcut2 * srcut2 * srcut2
        srcut12 = srcut6 * srcut6
        self.V_cut = 4.0 * epsilon * (srcut12 - srcut6)
        self.F_cut = 24.0 * epsilon * (2.0 * srcut12 - srcut6) / r_cut

    def params(self):
        return (self.epsilon, self.r_cut)


class LennardJonesXPLOR(Potential):
    """src/potentials.jl:176-182 (positional: eps, sigma, r_on, r_cut, tail_correction)"""
    tag = _capi.POT_LJ_XPLOR

    def __init__(self, epsilon, sigma, r_on, r_cut, tail_correction=False):
        self.epsilon, self.sigma, self.r_on, self.r_cut = float(epsilon), float(sigma), float(r_on), float(r_cut)
        self.tail_correction = tail_correction

    def params(self):
        return (self.epsilon, self.r_on, self.r_cut)


class Polydisperse(Potential):
    """The README's user-defined non-additive polydisperse plugin (README.md:82-145), shipped as a device functor."""
    tag = _capi.POT_POLY

    def __init__(self, rcut=1.25, non_additivity=0.2):
        self.rcut, self.non_additivity = float(rcut), float(non_additivity)

    def params(self):
        return (self.rcut, self.non_additivity)


class UserPotential(Potential):
    """User-defined interaction compiled for the device with NVRTC -- the open plugin contract of the reference
    (struct MyPot <: Potential + evaluate(::MyPot, r, sigma1, sigma2), src/types.jl:1-6, README.md:68-145).

    `body` is the CUDA-C body of
        __device__ bool evaluate(double r, double sigma1, double sigma2, const double* p, double& u, double& f)
    returning true when the pair interacts (false with u = f = 0 otherwise); `params` fills p[0..6]; `range` is the
    largest r at which it can return true.  Put the token MDB_DENSE_HITS in a comment of the body when most pairs within
    the range interact (e.g. Lennard-Jones): hits are then evaluated in line instead of through the deferred queue."""
    tag = _capi.POT_USER

    def __init__(self, body, params=(), range=1.0):
        self.body, self._params, self.range = str(body), tuple(float(v) for v in params), float(range)

    def params(self):
        return self._params


def energy_lrc(pot, n, volume):
    """src/potentials.jl:111-117,136-141,256-260,281-305: total long-range energy correction (0 unless enabled)."""
    rho = n / volume
    if isinstance(pot, LennardJones) and pot.tail_correction:
        sr = pot.sigma / pot.r_cut
        uij = ((sr ** 9) / 3.0) - (sr ** 3)
        uij *= 8.0 * math.pi * rho / 3.0
        return uij * n
    if isinstance(pot, LennardJonesXPLOR) and pot.tail_correction:
        s, e, rc = pot.sigma, pot.epsilon, pot.r_cut
        return (8.0 / 3.0) * math.pi * rho * n * e * s ** 3 * ((1.0 / 3.0) * (s / rc) ** 9 - (s / rc) ** 3)
    return 0.0


def pressure_lrc(pot, n, volume):
    """src/potentials.jl:123-128,149-152,267-271,291-313"""
    rho = n / volume
    if isinstance(pot, LennardJones) and pot.tail_correction:
        sr3 = (pot.sigma / pot.r_cut) ** 3
        result = (2.0 * sr3 ** 3 / 3.0) - sr3
        return result * 16.0 * math.pi * rho ** 2 / 3.0
    if isinstance(pot, LennardJonesXPLOR) and pot.tail_correction:
        s, e, rc = pot.sigma, pot.epsilon, pot.r_cut
        return (16.0 / 3.0) * math.pi * rho ** 2 * e * s ** 3 * ((2.0 / 3.0) * (s / rc) ** 9 - (s / rc) ** 3)
    return 0.0


# ---- src/types.jl:8-51 -------------------------------------------------------------------------------------
class Parameters:
    """Parameters(rho, n_particles, dt, potential)  (src/types.jl:8-13)"""

    def __init__(self, rho, n_particles, dt, potential):
        if not isinstance(potential, Potential):
            raise TypeError("potential must be a Potential subtype")
        self.rho = float(rho)
        self.n_particles = int(n_particles)
        self.dt = float(dt)
        self.potential = potential


class Ensemble:
    pass


class NVE(Ensemble):
    """src/types.jl:51"""


class NVT(Ensemble):
    """NVT(ktemp, tau); ktemp is a callable step -> T or a constant (src/types.jl:36-44)"""

    def __init__(self, ktemp, tau):
        self.ktemp = ktemp if callable(ktemp) else (lambda step, _t=float(ktemp): _t)
        self.tau = float(tau)


class Brownian(Ensemble):
    """src/types.jl:46-49"""

    def __init__(self, ktemp):
        self.ktemp = float(ktemp)


# ---- src/temperature_ramps.jl (scalar host callables feeding NVT.ktemp) -------------------------------------
class LinearRamp:
    def __init__(self, T_initial, T_final, n_steps):
        self.T_initial, self.T_final, self.n_steps = float(T_initial), float(T_final), int(n_steps)

    def __call__(self, step):
        if step > self.n_steps:
            return self.T_final
        step = min(max(step, 1), self.n_steps)
        if self.n_steps == 1:
            return self.T_final
        progress = (step - 1) / (self.n_steps - 1)
        return self.T_initial + (self.T_final - self.T_initial) * progress


class ExponentialRamp:
    def __init__(self, T_initial, T_final, n_steps):
        self.T_initial, self.T_final, self.n_steps = float(T_initial), float(T_final), int(n_steps)

    def __call__(self, step):
        if step > self.n_steps:
            return self.T_final
        step = min(max(step, 1), self.n_steps)
        if self.n_steps == 1 or self.T_initial == self.T_final:
            return self.T_final
        progress = (step - 1) / (self.n_steps - 1)
        return self.T_initial * math.exp(math.log(self.T_final / self.T_initial) * progress)


def initial_temperature_for_velocities(ktemp):
    """src/temperature_ramps.jl:67-73"""
    if hasattr(ktemp, "T_initial") and hasattr(ktemp, "T_final"):
        return max(ktemp.T_initial, ktemp.T_final)
    return ktemp


# ---- state ---------------------------------------------------------------------------------------------------
class EnergyAndForces:
    """View of the device-resident output of the pair map (src/types.jl:53-57)."""

    def __init__(self, engine):
        self._engine = engine

    @property
    def forces(self):
        return self._engine.download(positions=False, velocities=False, forces=True, images=False)[2]

    @property
    def energy(self):
        return float(self._engine.thermo()[0])

    @property
    def virial(self):
        return float(self._engine.thermo()[1])


class GPUSystem:
    """Stands where CellListMap.ParticleSystem stood in SimulationState.system (src/types.jl:15-17): exposes the same
    property names the loop touches (positions / xpositions / energy_and_forces) but keeps everything in HBM."""

    def __init__(self, engine, cutoff):
        self.engine = engine
        self.cutoff = cutoff
        self.energy_and_forces = EnergyAndForces(engine)

    @property
    def positions(self):
        return self.engine.download(positions=True, velocities=False, forces=False, images=False)[0]

    xpositions = positions

    def map_pairwise(self):
        """reset_output! + CellListMap.map_pairwise!(energy_and_forces!, system)  (src/simulation.jl:99-104)"""
        return self.engine.compute_forces()


class SimulationState:
    """src/types.jl:15-32 (same field names; arrays live on the GPU and are materialised on access)."""

    def __init__(self, system, diameters, rng, unitcell, dimension, nf):
        self.system = system
        self.diameters = diameters
        self.rng = rng
        self.unitcell = unitcell
        self.dimension = dimension
        self.nf = nf
        self._have_velocities = False

    @property
    def velocities(self):
        if not self._have_velocities:
            return np.zeros((0, self.dimension))  # empty, as initialize_state leaves it (src/initialization.jl:138)
        return self.system.engine.download(positions=False, velocities=True, forces=False, images=False)[1]

    @velocities.setter
    def velocities(self, v):
        self.system.engine.set_velocities(np.asarray(v, dtype=np.float64))
        self._have_velocities = True

    @property
    def images(self):
        return self.system.engine.download(positions=False, velocities=False, forces=False, images=True)[3]


def to_unitcell(box, dimension):
    """src/initialization.jl:7-18"""
    if np.isscalar(box):
        return float(box) * np.eye(dimension)
    box = np.asarray(box, dtype=np.float64)
    if box.ndim == 1:
        return np.diag(box[:dimension])
    if box.ndim == 2:
        return np.array(box[:dimension, :dimension])
    raise ValueError("Cannot interpret box/unitcell of type %s" % type(box))


def initialize_velocities(ktemp, rng, n_particles, dimension):
    """src/initialization.jl:32-47: randn(d, N), remove the centre-of-mass motion, rescale to ktemp.
    Returns an (N, d) array (the memory image of the reference's Vector{MVector})."""
    V = rng.standard_normal((dimension, n_particles))
    V -= V.mean(axis=1, keepdims=True)
    sum_v2 = float(np.sum(V * V))
    fs = math.sqrt(ktemp / (sum_v2 / ((n_particles - 1) * dimension)))
    V *= fs
    return np.ascontiguousarray(V.T)


def initialize_random(unitcell, npart, rng, dimension, tol=1.0, device=0, max_steps=400000, seed=None):
    """initialize_random(unitcell, npart, rng, dimension; tol=1.0)  (src/initialization.jl:20-30) on the GPU: uniform random
    points of the cell (mdb_random_positions), then the overlaps are removed -- the job Packmol's pack_monoatomic! does on
    the host in the reference -- by FIRE minimisation (mdb_fire_minimize) of the penalty u = k/2 (1 - r/tol')^2 on a
    scratch handle with MDB_POT_SOFT, tol' = 1.001 tol, until no pair is closer than tol (mdb_count_pairs).  Returns the
    (npart, dimension) positions; raises if the density does not admit such a packing within max_steps."""
    cell = np.asarray(to_unitcell(unitcell, dimension), dtype=np.float64)
    box = cell if np.any(cell != np.diag(np.diag(cell))) else np.diag(cell).copy()   # general cells go in as matrices
    if seed is None:
        seed = int(rng.integers(0, 2 ** 63 - 1))
    tol_pack = 1.001 * float(tol)
    eng = _capi.Engine(dimension, npart, box, tol_pack, _capi.POT_SOFT, (1.0, tol_pack), seed=seed, device=device)
    try:
        eng.upload(np.zeros((npart, dimension)), np.ones(npart), velocities=np.zeros((npart, dimension)))
        eng.random_positions(stream=0)
        done = 0
        while True:
            _, _, steps, _ = eng.fire_minimize(max_steps=2000, tol=1e-12, dt_initial=0.02, dt_max=0.2)
            done += steps
            close = eng.count_pairs(float(tol))
            if close == 0:
                break
            if done >= max_steps:
                raise _capi.MdbError(_capi.ERR_STATE, "initialize_random: %d pairs still closer than tol = %g after %d FIRE "
                                     "steps (density too high for this tolerance?)" % (close, tol, done))
        return eng.download(velocities=False, forces=False, images=False)[0]
    finally:
        eng.close()


def lattice_positions(n_particles, box, dimension, rng, jitter=0.02, sigma=1.0):
    """Deterministic overlap-free start (initialize_state(..., random_init="lattice")); random_init=True uses
    initialize_random like the reference.  Picks the lattice
    (sc / bcc / fcc, or sc / centred-rectangular in 2-D) with the largest nearest-neighbour distance that offers at least
    n_particles sites in the box, removes random sites down to n_particles and adds a small jitter that keeps
    neighbours further apart than `sigma` whenever the lattice allows it."""
    box = np.asarray(box, dtype=np.float64)[:dimension]
    bases = {"sc": np.zeros((1, dimension)), "centred": np.array([[0.0] * dimension, [0.5] * dimension])}
    if dimension == 3:
        bases["fcc"] = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    best = None
    for kind, basis in bases.items():
        nb = basis.shape[0]
        # cell counts proportional to the box edges, grown until there are enough sites
        scale = (n_particles / nb / np.prod(box)) ** (1.0 / dimension)
        m = np.maximum(1, np.floor(box * scale).astype(int))
        while nb * np.prod(m) < n_particles:
            m[np.argmax(box / m)] += 1
        a = box / m
        seps = [a[k] for k in range(dimension)]
        for bvec in basis[1:]:
            seps.append(float(np.sqrt(np.sum((bvec * a) ** 2))))
        nn = min(seps)
        if best is None or nn > best[0]:
            best = (nn, kind, m, basis)
    nn, kind, m, basis = best
    grid = np.stack(np.meshgrid(*[np.arange(mk) for mk in m], indexing="ij"), axis=-1).reshape(-1, dimension)
    frac = (grid[:, None, :] + basis[None, :, :] + 0.25).reshape(-1, dimension) / m
    keep = rng.permutation(frac.shape[0])[:n_particles]
    keep.sort()
    pos = frac[keep] * box
    jit = min(jitter, max(0.0, 0.45 * (nn - sigma) / math.sqrt(dimension)))
    pos += rng.uniform(-jit, jit, size=pos.shape) if jit > 0 else 0.0
    pos -= box * np.floor(pos / box)
    return pos


def write_to_file(filepath, step, unitcell, n_particles, positions, diameters, dimension, mode="a"):
    """extended-XYZ writer, src/io.jl:42-70"""
    with open(filepath, mode) as io:
        io.write("%d\n" % n_particles)
        flat = " ".join(repr(float(unitcell[i, j])) for j in range(dimension) for i in range(dimension))
        io.write('Lattice="%s" Properties=type:I:1:id:I:1:radius:R:1:pos:R:%d Time=%.6g\n' % (flat, dimension, step))
        for i in range(n_particles):
            io.write("%d %d %f" % (1, i + 1, diameters[i] / 2.0))
            for d in range(dimension):
                io.write(" %f" % positions[i, d])
            io.write("\n")


def read_file(filepath, dimension=3):
    """src/io.jl:176-205"""
    with open(filepath) as io:
        n = int(io.readline())
        header = io.readline()
        m = re.search(r'Lattice="([^"]+)"', header)
        if not m:
            raise ValueError("Could not parse Lattice property in file header")
        entries = [float(t) for t in m.group(1).split()]
        unitcell = np.array(entries).reshape(dimension, dimension).T
        pos = np.zeros((n, dimension))
        radii = np.zeros(n)
        for i in range(n):
            line = io.readline().split(" ")
            radii[i] = float(line[2])
            pos[i] = [float(t) for t in line[3:3 + dimension]]
    return unitcell, pos, radii * 2.0


def write_to_file_lammps(filepath, step, unitcell, n_particles, positions, images, diameters, dimension, mode="w"):
    """LAMMPS dump with wrapped + unwrapped coordinates, src/io.jl:96-170 (unwrapped = p + U*img, :78-86)"""
    boxmat = np.zeros((3, 3))
    boxmat[:dimension, :dimension] = unitcell
    uw = positions + images @ np.asarray(unitcell).T
    with open(filepath, mode) as io:
        io.write("ITEM: TIMESTEP\n%d\n" % step)
        io.write("ITEM: NUMBER OF ATOMS\n%d\n" % n_particles)
        if dimension == 2:
            lx, ly = np.linalg.norm(boxmat[:, 0]), np.linalg.norm(boxmat[:, 1])
            io.write("ITEM: BOX BOUNDS xy pp pp\n")
            io.write("%f %f %f\n" % (0.0, lx, boxmat[0, 1]))
            io.write("%f %f 0.0\n" % (0.0, ly))
            io.write("%f %f 0.0\n" % (0.0, 1.0))
            io.write("ITEM: ATOMS id type radius x y xu yu\n")
        elif dimension == 3:
            io.write("ITEM: BOX BOUNDS xy xz yz pp pp pp\n")
            io.write("%f %f %f\n" % (0.0, np.linalg.norm(boxmat[:, 0]), boxmat[0, 1]))
            io.write("%f %f %f\n" % (0.0, np.linalg.norm(boxmat[:, 1]), boxmat[1, 2]))
            io.write("%f %f %f\n" % (0.0, np.linalg.norm(boxmat[:, 2]), boxmat[0, 2]))
            io.write("ITEM: ATOMS id type radius x y z xu yu zu\n")
        else:
            raise ValueError("Unsupported dimension: %d" % dimension)
        cols = np.column_stack([np.arange(1, n_particles + 1), np.ones(n_particles), diameters / 2.0, positions, uw])
        fmt = "%d %d " + " ".join(["%f"] * (1 + 2 * dimension))
        np.savetxt(io, cols, fmt=fmt)


def initialize_state(params, pathname, from_file="", dimension=3, random_init=False, cutoff=1.5, rng=None, unitcell=None,
                     positions=None, diameters=None, device=0, mode="auto", skin=None, use_graph=True, seed=None,
                     write_init=True):
    """src/initialization.jl:112-157 (+ initialize_simulation :49-110).  Extra keyword arguments (device, mode, skin,
    use_graph, seed) configure the GPU engine and have no counterpart in the reference."""
    if rng is None:
        rng = np.random.default_rng()
    n_particles = params.n_particles
    nf = dimension * (params.n_particles - 1.0)
    if positions is not None and diameters is not None:
        positions = np.asarray(positions, dtype=np.float64)
        n_particles = positions.shape[0]
        if unitcell is None:
            box_vec = positions.max(axis=0) - positions.min(axis=0)
            unitcell = to_unitcell(box_vec, dimension)
        else:
            unitcell = to_unitcell(unitcell, dimension)
    elif os.path.isfile(from_file) or not random_init:
        unitcell, positions, diameters = read_file(from_file, dimension=dimension)
        n_particles = positions.shape[0]
    else:
        # initialize_simulation's random branch (src/initialization.jl:86-92): box from the density unless a cell is given
        if unitcell is not None:
            unitcell = to_unitcell(unitcell, dimension)
        else:
            unitcell = to_unitcell((n_particles / params.rho) ** (1.0 / dimension), dimension)
        if random_init == "lattice":
            positions = lattice_positions(n_particles, np.diag(unitcell), dimension, rng)
        else:
            positions = initialize_random(unitcell, n_particles, rng, dimension, device=device)
        diameters = np.ones(n_particles)
    diameters = np.ascontiguousarray(diameters, dtype=np.float64)
    triclinic = bool(np.any(unitcell != np.diag(np.diag(unitcell))))   # full matrix: lattice vectors in the columns
    pot = params.potential
    if pot.tag is None:
        # no device functor: mirror the reference's `error("evaluate not implemented ...")` (src/types.jl:4-6)
        evaluate(pot, 1.0)
    modes = {"auto": _capi.MODE_AUTO, "cells": _capi.MODE_CELLS, "list": _capi.MODE_LIST, "small": _capi.MODE_SMALL}
    if seed is None:
        seed = int(rng.integers(0, 2 ** 63 - 1))
    user = isinstance(pot, UserPotential)
    engine = _capi.Engine(dimension, n_particles, np.asarray(unitcell) if triclinic else np.diag(unitcell), float(cutoff), _capi.POT_PSEUDOHS if user else pot.tag,
                          () if user else pot.params(), seed=seed, device=device, mode=modes[mode], skin=skin or 0.0,
                          use_graph=use_graph)
    if user:
        engine.set_user_potential(pot.body, pot.params(), pot.range)
    engine.upload(positions, diameters)
    system = GPUSystem(engine, cutoff)
    state = SimulationState(system, diameters, rng, unitcell, dimension, nf)
    if write_init and pathname is not None:
        write_to_file(os.path.join(pathname, "init.xyz"), 0, unitcell, n_particles, system.positions, diameters,
                      dimension, mode="w")
    return state


def compute_box_volume(unitcell):
    """src/simulation.jl:7-9"""
    return abs(float(np.linalg.det(unitcell)))


def _open_files(pathname, traj_name, thermo_name):
    """src/io.jl:225-239"""
    traj, thermo = os.path.join(pathname, traj_name), os.path.join(pathname, thermo_name)
    for f in (traj, thermo):
        if os.path.isfile(f):
            os.remove(f)
    return traj, thermo


def generate_log_times(max_iter=10000, logn=40, logbase=1.35):
    """src/io.jl:17-36 (without the side-effect file)"""
    maxlog = int(math.floor(logbase ** logn))
    times = set()
    for j in range(max_iter + 1):
        for i in range(logn + 1):
            times.add(int(math.floor(j * maxlog + logbase ** i)))
    return sorted(times)


def run_simulation(state, params, ensemble, total_steps, frequency, pathname, traj_name="trajectory.xyz",
                   thermo_name="thermo.txt", compress=False, log_times=False, write_trajectory=True):
    """run_simulation!(state, params, ensemble, total_steps, frequency, pathname; ...)  (src/simulation.jl:40-178 for
    NVE/NVT, :181-308 for Brownian).  The step loop runs on the GPU in chunks that end at the steps where the reference
    writes output; rows and frames are produced from the device state exactly where the reference produces them.
    Returns the per-step thermo array [total_steps][4] = (U, W, KE, n_pairs) in addition to writing the files."""
    engine = state.system.engine
    traj_file, thermo_file = _open_files(pathname, traj_name, thermo_name)
    with open(thermo_file, "a") as io:
        io.write("# Step Energy Temperature Pressure\n")
    dimension = state.dimension
    potential = params.potential
    volume = compute_box_volume(state.unitcell)
    n = params.n_particles
    is_bd = isinstance(ensemble, Brownian)
    if not is_bd and not state._have_velocities:
        raise _capi.MdbError(_capi.ERR_STATE, "state.velocities was never set (README.md:38-41)")
    snapshot_times = None
    if log_times:
        snapshot_times = [0] + generate_log_times()
        snap_index = 0
    # chunk boundaries: after every step with step % frequency == 0, and after log-time steps
    stops = set(range(0, total_steps, frequency)) if frequency > 0 else set()
    if log_times:
        stops |= {t for t in snapshot_times if t < total_steps}
    stops = sorted(stops)
    all_thermo = np.zeros((total_steps, 4))
    done = 0
    virial_acc, nprom = 0.0, 0

    def advance(upto):  # run steps done .. upto-1
        nonlocal done, virial_acc, nprom
        m = upto - done
        if m <= 0:
            return
        if isinstance(ensemble, NVT):
            kt = np.array([ensemble.ktemp(s + 1) for s in range(done, upto)], dtype=np.float64)
            t = engine.run_nvt(m, params.dt, kt, ensemble.tau)
        elif is_bd:
            t = engine.run_brownian(m, params.dt, ensemble.ktemp)
            for s in range(done, upto):  # virial sampled every 10 steps (src/simulation.jl:253-256)
                if s % 10 == 0:
                    virial_acc += t[s - done, 1]
                    nprom += 1
        else:
            t = engine.run_nve(m, params.dt)
        all_thermo[done:upto] = t
        done = upto

    frame_slot = 0

    def emit_frame(path, step, append):
        # write_to_file_lammps (src/io.jl:78-170): unwrapped coordinates formed and the frame packed on the device, copied
        # on the engine's copy stream and formatted by its writer thread while the next chunk of steps runs
        nonlocal frame_slot
        engine.frame_capture(frame_slot)
        engine.frame_write_lammps(frame_slot, path, step, append=append)
        frame_slot = (frame_slot + 1) % _capi.FRAME_SLOTS

    for step in stops:
        advance(step + 1)
        U, W, KE, _ = all_thermo[step]
        if step % frequency == 0:
            if is_bd:
                ener = U / n
                # nprom == 0 (no step with step % 10 == 0 in this output interval): the reference evaluates 0.0/0 = NaN,
                # writes the row and carries on (src/simulation.jl:252-262)
                pressure = (virial_acc / (dimension * nprom * volume) if nprom else float("nan")) + params.rho * ensemble.ktemp
                row = (step, ener, ensemble.ktemp, pressure)
                virial_acc, nprom = 0.0, 0
            else:
                total_energy = (U + energy_lrc(potential, n, volume)) / n
                temperature = 2.0 * KE / state.nf
                pressure = W / (dimension * volume) + params.rho * temperature
                pressure += pressure_lrc(potential, n, volume)
                row = (step, total_energy, temperature, pressure)
            with open(thermo_file, "a") as io:
                io.write("%d %.6f %.6f %.6f\n" % row)
            if write_trajectory:
                emit_frame(traj_file, step, True)
        if log_times and snap_index < len(snapshot_times) and snapshot_times[snap_index] == step:
            emit_frame(os.path.join(pathname, "snapshot.%d" % step), step, False)
            snap_index += 1
    advance(total_steps)
    engine.frame_flush()
    # finalize_simulation! (src/simulation.jl:11-36)
    write_to_file(os.path.join(pathname, "final.xyz"), total_steps, state.unitcell, n, state.system.positions,
                  state.diameters, dimension, mode="w")
    if compress and os.path.isfile(traj_file):
        try:
            import zstandard
            with open(traj_file, "rb") as fi, open(traj_file + ".zst", "wb") as fo:
                fo.write(zstandard.ZstdCompressor().compress(fi.read()))
            os.remove(traj_file)
        except ImportError:
            pass  # zstd codec absent in this image: trajectory is left uncompressed (I/O is outside the hot path)
    return all_thermo


def minimize(state, params, pathname, dimension, method="FIRE", save_config="minimized.xyz", **kwargs):
    """minimize!(state, params, pathname, dimension; method=:FIRE, ...)  (src/minimize.jl:166-197)"""
    if method != "FIRE":
        raise ValueError("Unknown minimization method: %s" % method)
    energy, frms, steps, converged = state.system.engine.fire_minimize(**kwargs)
    if pathname is not None:
        write_to_file(os.path.join(pathname, save_config), 0, state.unitcell, params.n_particles,
                      state.system.positions, state.diameters, dimension, mode="w")
    return (energy, converged) if converged else None
