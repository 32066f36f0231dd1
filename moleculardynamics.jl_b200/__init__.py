"""B200-native engine for MolecularDynamics.jl's per-step hot path (pair forces + integrators + thermostat).

Importable as `mdjl_b200` (see ../mdjl_b200.py; the directory name carries a dot).  The public names mirror the
reference's exports (src/MolecularDynamics.jl:29-35).
"""
from . import _build, _capi
from ._capi import Engine, MdbError, SlabRing, unique_id
from .api import (NVE, NVT, Brownian, EnergyAndForces, ExponentialRamp, GPUSystem, LennardJones, LennardJonesXPLOR,
                  LinearRamp, Parameters, Polydisperse, Potential, PseudoHS, SimulationState, UserPotential, energy_lrc, evaluate,
                  initial_temperature_for_velocities, initialize_random, initialize_state, initialize_velocities, lattice_positions,
                  minimize, pressure_lrc, read_file, run_simulation, to_unitcell, write_to_file, write_to_file_lammps)

__all__ = [
    "Parameters", "NVT", "NVE", "Brownian", "initialize_state", "run_simulation", "PseudoHS", "LennardJonesXPLOR",
    "LennardJones", "Polydisperse", "UserPotential", "LinearRamp", "ExponentialRamp", "minimize", "initial_temperature_for_velocities",
    "initialize_velocities", "initialize_random", "Potential", "evaluate", "Engine", "MdbError",
]
