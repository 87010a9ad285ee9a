"""Drop-in for the reference's `MCMC` package (hot-path subset).

Mirrors MCMC/__init__.py:11-28 of the reference for the classes and functions
on the sampling path and the lattice initialisers (MCMC/initialise.py, without their plots); plotting and the
single-chain CLI are out of scope.
"""
from .simulation_box import SimulationBox
from .potential import lennard_jones_energy_virial, double_well_potential
from .energy_calculator import EnergyCalculator
from .batched import BatchedMonteCarlo
from .monte_carlo import MonteCarlo
from .initialise import (initialise_low_left, initialise_low_right, initialise_fcc, initialise_cluster,
                         initialise_chains, jittered_lattice)

__all__ = ["SimulationBox", "EnergyCalculator", "MonteCarlo", "BatchedMonteCarlo",
           "lennard_jones_energy_virial", "double_well_potential",
           "initialise_low_left", "initialise_low_right", "initialise_fcc", "initialise_cluster", "initialise_chains",
           "jittered_lattice"]
