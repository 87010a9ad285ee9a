"""EnergyCalculator drop-in.  Mirrors MCMC/energy_calculator.py:10-203 of the
reference: same constructor, `total_energy` / `total_virial` attributes, and the
three methods; evaluation goes to fs_energy_total / fs_energy_particle."""
import time

import numpy as np
import torch

from ._bridge import _lib
from .simulation_box import _dev


class EnergyCalculator:
    def __init__(self, num_particles, initial_particles, simulation_box, num_wells=0, V0_list=[-4.0, -4.2],
                 r0=1.0, k=10.0, timing=True, checking=False):
        self.sim_box = simulation_box
        self.num_particles = num_particles
        self.num_wells = num_wells
        self.V0_list = V0_list
        self.r0 = r0
        self.k = k
        self.timing = timing          # accepted and inert: the reference prints per call
        self.checking = checking
        self.particle_energy_times = []
        self.total_energy_times = []
        self._pot = _lib.make_pot(num_wells, V0_list, r0, k)
        self._dev = _dev()
        _lib.lib()
        self.total_energy, self.total_virial = self.calculate_total_energy_virial(initial_particles)

    def _upload(self, positions):
        if torch.is_tensor(positions):
            t = positions
        else:
            t = torch.as_tensor(np.ascontiguousarray(positions))
        t = t.to(self._dev, torch.float32).reshape(-1, self.num_particles, 2).contiguous()
        return t

    def _box(self):
        return float(self.sim_box.box_size_x), float(self.sim_box.box_size_y)

    def calculate_particle_energy_virial(self, positions, particle_index):
        """:48-108 - (energy, virial) of one particle; (inf, inf) on a hard-core overlap."""
        t0 = time.time()
        pos = self._upload(positions)
        idx = torch.tensor([int(particle_index)], dtype=torch.int32, device=self._dev)
        e = torch.empty(1, dtype=torch.float32, device=self._dev)
        w = torch.empty(1, dtype=torch.float32, device=self._dev)
        Lx, Ly = self._box()
        _lib.check(_lib.lib().fs_energy_particle(_lib.ptr(pos), _lib.ptr(idx), None, 1, self.num_particles, Lx, Ly,
                                                 self._pot, _lib.ptr(e), _lib.ptr(w), None, _lib.stream_ptr(self._dev)))
        out = float(e.item()), float(w.item())
        self.particle_energy_times.append(time.time() - t0)
        return out

    def update_total_energy_virial(self, energy_dif, virial_dif):
        """:110-119"""
        self.total_energy += energy_dif
        self.total_virial += virial_dif

    def calculate_total_energy_virial(self, positions):
        """:121-203 - overwrites the cached totals like the reference (:135-136, 189)."""
        t0 = time.time()
        pos = self._upload(positions)
        E = torch.empty(1, dtype=torch.float32, device=self._dev)
        W = torch.empty(1, dtype=torch.float32, device=self._dev)
        Lx, Ly = self._box()
        _lib.check(_lib.lib().fs_energy_total(_lib.ptr(pos), 1, self.num_particles, Lx, Ly, self._pot,
                                              _lib.ptr(E), _lib.ptr(W), None, _lib.stream_ptr(self._dev)))
        self.total_energy, self.total_virial = float(E.item()), float(W.item())
        self.total_energy_times.append(time.time() - t0)
        return self.total_energy, self.total_virial
