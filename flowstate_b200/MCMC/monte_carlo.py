"""MonteCarlo drop-in: the reference's single-chain sampler API
(MCMC/monte_carlo.py:11-475) as a B = 1 view of BatchedMonteCarlo.

Same constructor signature, attributes read by the drivers (`particles`,
`max_displacement`, `attempts_displacement`, `accepted_displacement`,
`energy_calculator`, `local_samples`, `testing_samples`, `logger`, `rng`,
`nf_model`, `device`) and methods.  `rng` stays a numpy Generator whose PCG64
state is mirrored to the device before and after every kernel call, so a chain
seeded like the reference draws the same numbers in the same order.
"""
import numpy as np
import torch

from ._bridge import _lib
from .batched import BatchedMonteCarlo, pcg64_set_state, pcg64_state_words


class _EnergyView:
    """`mc.energy_calculator` of the reference: cached totals + the three methods,
    backed by the chain's device state."""

    def __init__(self, mc):
        self._mc = mc
        self.sim_box = mc.sim_box
        self.num_particles = mc.num_particles
        self.particle_energy_times = []
        self.total_energy_times = []

    @property
    def total_energy(self):
        return float(self._mc._b.E.item())

    @total_energy.setter
    def total_energy(self, v):
        self._mc._b.E.fill_(float(v))

    @property
    def total_virial(self):
        return float(self._mc._b.W.item())

    @total_virial.setter
    def total_virial(self, v):
        self._mc._b.W.fill_(float(v))

    def calculate_total_energy_virial(self, positions):
        b = self._mc._b
        cfg = torch.as_tensor(np.ascontiguousarray(positions)).to(b.device, torch.float32).reshape(1, -1, 2)
        E, W, _ = b.total_energy_virial(cfg.contiguous())
        self.total_energy, self.total_virial = float(E.item()), float(W.item())
        return self.total_energy, self.total_virial

    def calculate_particle_energy_virial(self, positions, particle_index):
        b = self._mc._b
        pos = torch.as_tensor(np.ascontiguousarray(positions)).to(b.device, torch.float32).reshape(1, -1, 2).contiguous()
        idx = torch.tensor([int(particle_index)], dtype=torch.int32, device=b.device)
        e = torch.empty(1, dtype=torch.float32, device=b.device)
        w = torch.empty(1, dtype=torch.float32, device=b.device)
        Lx, Ly = b._L()
        _lib.check(_lib.lib().fs_energy_particle(_lib.ptr(pos), _lib.ptr(idx), None, 1, b.num_particles, Lx, Ly,
                                                 b._pot, _lib.ptr(e), _lib.ptr(w), None, _lib.stream_ptr(b.device)))
        return float(e.item()), float(w.item())

    def update_total_energy_virial(self, energy_dif, virial_dif):
        self.total_energy = self.total_energy + energy_dif
        self.total_virial = self.total_virial + virial_dif


class MonteCarlo:
    def __init__(self, particles, sim_box, temperature, num_particles, num_wells=0, V0_list=[-0.5, -0.5], r0=1.0,
                 k=10, initial_max_displacement=0.5, target_acceptance=0.5, timing=False, checking=False,
                 logger=None, seed=None, device=None):
        self.sim_box = sim_box
        self.half_width = sim_box.box_size_x / 2
        self.beta = 1.0 / temperature
        self.num_particles = num_particles
        self.num_wells, self.V0_list, self.r0, self.k = num_wells, V0_list, r0, k
        self.target_acceptance = target_acceptance
        self.timing, self.checking, self.debug = timing, checking, True
        self.logger = logger
        self.rng = np.random.default_rng(seed=seed) if seed is not None else np.random.default_rng()
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.device = torch.device(device)
        self._b = BatchedMonteCarlo(np.asarray(particles)[None], sim_box, temperature, num_particles,
                                    num_wells=num_wells, V0_list=V0_list, r0=r0, k=k,
                                    initial_max_displacement=initial_max_displacement,
                                    target_acceptance=target_acceptance, seeds=[0], device=self.device)
        self.energy_calculator = _EnergyView(self)
        self.pressure_history = []
        self.volume_history = [self.sim_box.volume]
        self.densities = []
        self.running_mean_window = 1000
        self.particle_move_times = []
        self.last_configuration = ["left" if p[0] < self.half_width else "right" for p in self.particles]
        self.local_samples = []
        self.testing_samples = []
        self.nf_model = None

    # -- state views ------------------------------------------------------
    @property
    def particles(self):
        return self._b.pos[0].cpu().numpy()

    @particles.setter
    def particles(self, value):
        self._b.pos[0].copy_(torch.as_tensor(np.ascontiguousarray(value)).to(self._b.device, torch.float32))

    @property
    def max_displacement(self):
        return float(self._b.max_disp.item())

    @max_displacement.setter
    def max_displacement(self, v):
        self._b.max_disp.fill_(float(v))

    def _counter(name):   # noqa: N805
        def get(self):
            return int(getattr(self._b, name).item())

        def set_(self, v):
            getattr(self._b, name).fill_(int(v))
        return property(get, set_)

    attempts_displacement = _counter("attempts")
    accepted_displacement = _counter("accepted")
    previous_attempts_displacement = _counter("prev_attempts")
    previous_accepted_displacement = _counter("prev_accepted")
    del _counter

    def _push_rng(self):
        w = np.array([pcg64_state_words(self.rng)], dtype=np.uint64)
        self._b.pcg_state.copy_(torch.from_numpy(w.view(np.int64)))

    def _pull_rng(self):
        w = self._b.pcg_state.cpu().numpy().view(np.uint64)[0]
        pcg64_set_state(self.rng, w)

    def _log(self, message, level="info"):
        if self.logger:
            getattr(self.logger, level if level in ("debug", "warning", "error") else "info")(message)
        else:
            print(message)

    # -- moves ------------------------------------------------------------
    def particle_displacement(self):
        """:146-189"""
        self._push_rng()
        tr = self._b.particle_displacement(1, trace=True)
        self._pull_rng()
        if int(tr["accept"].item()):
            p = int(tr["idx"].item())
            side = "left" if self.particles[p][0] < self.half_width else "right"
            if side != self.last_configuration[p]:
                self._log("Particle %d has crossed from %s to %s." % (p, self.last_configuration[p], side))
                self.last_configuration[p] = side

    def metropolis_acceptance_particle_move(self, old_energy, new_energy):
        """:191-223 (host rule for callers that bring their own energies)."""
        if new_energy <= old_energy:
            return True
        if np.isinf(new_energy):
            return False
        return self.rng.random() < np.exp(-self.beta * (new_energy - old_energy))

    def set_nf_model(self, nf_model):
        self.nf_model = nf_model
        self._b.set_nf_model(nf_model)

    def nf_big_move(self, config):
        """:235-303"""
        cfg = torch.as_tensor(np.ascontiguousarray(config)).to(self._b.device, torch.float32).reshape(1, -1, 2)
        self._push_rng()
        mask = self._b.nf_big_move(cfg.contiguous())
        self._pull_rng()
        return bool(mask.item())

    def judge_normalizing_flow(self, config):
        """:305-330 energy-only Metropolis test; the cached energy is left untouched."""
        self.attempts_displacement += 1
        eno, viro = self.energy_calculator.total_energy, self.energy_calculator.total_virial
        enn, _ = self.energy_calculator.calculate_total_energy_virial(config)
        crit = self.metropolis_acceptance_particle_move(eno, enn)
        self.energy_calculator.total_energy, self.energy_calculator.total_virial = eno, viro
        return crit

    def bulk_judge_normalizing_flow(self, configs, ref_energy):
        """:332-370"""
        b = self._b
        cfg = torch.as_tensor(np.ascontiguousarray(np.asarray(configs))).to(b.device, torch.float32)
        E, _, _ = b.total_energy_virial(cfg.reshape(-1, b.num_particles, 2).contiguous())
        accepted = sum(1 for e in E.cpu().numpy().astype(np.float64)
                       if self.metropolis_acceptance_particle_move(ref_energy, e))
        self._log("Bulk judge normalizing flow: %d accepted moves out of %d attempted moves (reference energy: %.3f)."
                  % (accepted, len(configs), ref_energy))
        return accepted, len(configs)

    def adjust_displacement(self):
        """:375-403"""
        self._b.adjust_displacement()

    def adjust_volume(self):
        return

    def sample(self, cycle_number):
        """:416-444"""
        energy_per_particle = self.energy_calculator.total_energy / self.num_particles
        volume = self.sim_box.volume
        density = self.num_particles / volume
        pressure = density / self.beta + self.energy_calculator.total_virial / (2.0 * volume)
        self.pressure_history.append(pressure)
        self.volume_history.append(volume)
        self.densities.append(density)
        return (cycle_number, energy_per_particle, density, pressure, self.sim_box.box_size_x,
                self.sim_box.box_size_y, self.particles.copy())

    def check_equilibration(self, tolerance=0.05, window=500):
        """:449-475"""
        if len(self.pressure_history) < window:
            return False
        p = np.asarray(self.pressure_history[-window:])
        d = np.asarray(self.densities[-window:])
        conds = [(p.std() / p.mean() < tolerance) if p.mean() != 0 else False,
                 (d.std() / d.mean() < tolerance) if d.mean() != 0 else False]
        return all(conds)
