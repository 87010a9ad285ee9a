"""BatchedMonteCarlo: B independent Metropolis chains resident on one GPU.

The reference runs "parallel chains" as a Python loop over MonteCarlo objects
(hybrid_NF_MCMC/main_algorithm_1.py:203-210, 381-395); here the same chain
logic (MCMC/monte_carlo.py:146-303, 375-444) is vectorised over chains and
executed by the CUDA kernels fs_local_sweep / fs_energy_total /
fs_accept_global / fs_adjust_displacement.  State per chain keeps the
reference's types: float32 positions (what a reference chain holds after its
first NF acceptance), float64 running energy/virial and max displacement,
int64 counters.
"""
import ctypes as C

import numpy as np
import torch

from ._bridge import _lib
from .simulation_box import SimulationBox, _dev


def pcg64_state_words(rng):
    """numpy Generator(PCG64) -> 6 uint64 words {state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger}."""
    st = rng.bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise _lib.FlowStateError("flowstate_b200: only numpy PCG64 generators can be mirrored on the device")
    s, inc = st["state"]["state"], st["state"]["inc"]
    m = (1 << 64) - 1
    return [s >> 64, s & m, inc >> 64, inc & m, st["has_uint32"], st["uinteger"]]


def pcg64_set_state(rng, words):
    w = [int(x) & ((1 << 64) - 1) for x in words]
    rng.bit_generator.state = {"bit_generator": "PCG64",
                               "state": {"state": (w[0] << 64) | w[1], "inc": (w[2] << 64) | w[3]},
                               "has_uint32": int(w[4]), "uinteger": int(w[5])}


class BatchedMonteCarlo:
    """B chains of `num_particles` particles in one box.

    particles: (B, N, 2) array/tensor in MC-box coordinates.
    seeds:     iterable of B ints -> per-chain numpy-compatible PCG64 streams
               (chain i reproduces np.random.default_rng(seeds[i]), the reference's
               per-chain generator, monte_carlo.py:92-95); or None with
               rng="philox" for counter-based streams keyed by the global chain id
               (the throughput kernel: several chains per warp).
    """

    def __init__(self, particles, sim_box, temperature, num_particles, num_wells=0, V0_list=(-0.5, -0.5),
                 r0=1.0, k=10, initial_max_displacement=0.5, target_acceptance=0.5, seeds=None, rng="pcg64",
                 philox_seed=0, chain_id0=0, device=None):
        self.device = torch.device(device) if device is not None else _dev()
        if self.device.type != "cuda":
            raise _lib.FlowStateError("flowstate_b200: BatchedMonteCarlo needs a CUDA device (no CPU fallback)")
        _lib.lib()
        if isinstance(sim_box, (int, float)):
            sim_box = SimulationBox(sim_box)
        self.sim_box = sim_box
        self.half_width = sim_box.box_size_x / 2
        self.beta = 1.0 / temperature
        self.num_particles = int(num_particles)
        self.num_wells, self.V0_list, self.r0, self.k = num_wells, list(V0_list), r0, k
        self.target_acceptance = target_acceptance
        self._pot = _lib.make_pot(num_wells, V0_list, r0, k)
        pos = torch.as_tensor(np.asarray(particles) if not torch.is_tensor(particles) else particles)
        if pos.dim() == 2:
            pos = pos[None]
        if pos.shape[1:] != (self.num_particles, 2):
            raise ValueError("particles must have shape (B, %d, 2), got %s" % (self.num_particles, tuple(pos.shape)))
        self.pos = pos.to(self.device, torch.float32).contiguous().clone()
        B = self.B = self.pos.shape[0]
        dev = self.device
        self.E = torch.zeros(B, dtype=torch.float64, device=dev)
        self.W = torch.zeros(B, dtype=torch.float64, device=dev)
        self.max_disp = torch.full((B,), float(initial_max_displacement), dtype=torch.float64, device=dev)
        self.attempts = torch.zeros(B, dtype=torch.int64, device=dev)
        self.accepted = torch.zeros(B, dtype=torch.int64, device=dev)
        self.prev_attempts = torch.zeros(B, dtype=torch.int64, device=dev)
        self.prev_accepted = torch.zeros(B, dtype=torch.int64, device=dev)
        self.rng_kind = rng
        self.philox_seed = int(philox_seed)
        self.chain_id0 = int(chain_id0)
        self.pcg_state = None
        if rng == "pcg64":
            if seeds is None:
                seeds = [None] * B
            words = np.array([pcg64_state_words(np.random.default_rng(s)) for s in seeds], dtype=np.uint64)
            self.pcg_state = torch.from_numpy(words.view(np.int64)).to(dev).contiguous()
        elif rng not in ("philox", "philox_ref"):
            raise ValueError("rng must be 'pcg64', 'philox' or 'philox_ref'")
        self.nf_model = None
        self.launches = 0          # kernels launched through this object (bench bookkeeping)
        self.fused_accept = True   # nf_big_move: proposal energy + acceptance + update in one kernel
        self.proposal_energy = None
        self.refresh_energy()

    # -- plumbing ---------------------------------------------------------
    def _rng_struct(self, replay=None):
        r = _lib.FsRng()
        if replay is not None:
            r.kind = _lib.FS_RNG_REPLAY
            idx, u, cursor = replay
            r.replay_idx, r.replay_u, r.replay_cursor = idx.data_ptr(), u.data_ptr(), cursor.data_ptr()
            r.idx_stride, r.u_stride = idx.shape[1], u.shape[1]
        elif self.rng_kind == "pcg64":
            r.kind = _lib.FS_RNG_PCG64
            r.pcg_state = self.pcg_state.data_ptr()
        else:
            # "philox_ref": the same counter-based draws through the reference-order parity kernel (tests)
            r.kind = _lib.FS_RNG_PHILOX if self.rng_kind == "philox" else _lib.FS_RNG_PHILOX_REF
            r.philox_seed = self.philox_seed
            r.chain_id0 = self.chain_id0
        return r

    def _L(self):
        return float(self.sim_box.box_size_x), float(self.sim_box.box_size_y)

    # -- energies ---------------------------------------------------------
    def total_energy_virial(self, configs=None):
        """EnergyCalculator.calculate_total_energy_virial for every chain (or for `configs`
        (B', N, 2)); returns float32 tensors (E, W, overlap)."""
        pos = self.pos if configs is None else _lib.require_cuda(configs, "configs")
        if pos.dtype != torch.float32:
            raise _lib.FlowStateError("flowstate_b200: configurations must be float32")
        Bc = pos.shape[0]
        E = torch.empty(Bc, dtype=torch.float32, device=self.device)
        W = torch.empty(Bc, dtype=torch.float32, device=self.device)
        ov = torch.empty(Bc, dtype=torch.uint8, device=self.device)
        Lx, Ly = self._L()
        _lib.check(_lib.lib().fs_energy_total(_lib.ptr(pos), Bc, self.num_particles, Lx, Ly, self._pot,
                                              _lib.ptr(E), _lib.ptr(W), _lib.ptr(ov), _lib.stream_ptr(self.device)))
        self.launches += 1
        return E, W, ov

    def refresh_energy(self):
        """Recomputes the running totals from the positions (drops incremental drift)."""
        E, W, _ = self.total_energy_virial()
        self.E.copy_(E.double())
        self.W.copy_(W.double())

    # -- local moves ------------------------------------------------------
    def particle_displacement(self, steps=1, trace=False, replay=None):
        """`steps` local Metropolis moves on every chain (monte_carlo.py:146-223).
        trace=True returns dict(accept uint8 [B,steps], idx int32, e float32 [B,steps,2])."""
        B, dev = self.B, self.device
        ta = ti = te = None
        if trace:
            ta = torch.empty(B, steps, dtype=torch.uint8, device=dev)
            ti = torch.empty(B, steps, dtype=torch.int32, device=dev)
            te = torch.empty(B, steps, 2, dtype=torch.float32, device=dev)
        Lx, Ly = self._L()
        rng = self._rng_struct(replay)
        _lib.check(_lib.lib().fs_local_sweep(
            _lib.ptr(self.pos), _lib.ptr(self.E), _lib.ptr(self.W), _lib.ptr(self.max_disp),
            _lib.ptr(self.attempts), _lib.ptr(self.accepted), B, self.num_particles, int(steps), Lx, Ly,
            float(self.beta), self._pot, C.byref(rng), _lib.ptr(ta), _lib.ptr(ti), _lib.ptr(te),
            _lib.stream_ptr(dev)))
        self.launches += 1
        if trace:
            return {"accept": ta, "idx": ti, "e": te}
        return None

    def adjust_displacement(self):
        """monte_carlo.py:375-403 for every chain."""
        _lib.check(_lib.lib().fs_adjust_displacement(
            _lib.ptr(self.max_disp), _lib.ptr(self.attempts), _lib.ptr(self.accepted),
            _lib.ptr(self.prev_attempts), _lib.ptr(self.prev_accepted), float(self.target_acceptance), self.B,
            _lib.stream_ptr(self.device)))
        self.launches += 1

    # -- global moves -----------------------------------------------------
    def set_nf_model(self, nf_model):
        self.nf_model = nf_model

    def centred(self, pos):
        """MC-box -> flow coordinates, float64 subtraction stored float32 (monte_carlo.py:251-258)."""
        return (pos.double() - self.half_width).float().reshape(pos.shape[0], -1)

    def nf_big_move(self, configs=None, u=None, logq=None, logq_new=None):
        """One NF-proposed global move per chain (monte_carlo.py:235-303).

        configs: (B, N, 2) float32 CUDA tensor in MC-box coordinates; None draws
                 `nf_model.sample(B) + half_width` (main_algorithm_2.py:479-482).
        u:       optional float64 [B] uniforms (replay); default draws from the chain RNG.
        logq:    optional (logq_old, logq_new) float32 tensors to skip the flow.
        logq_new: optional float32 [B] log q of the proposals when the caller already holds it (the sampling pass that
                 produced them yields it: NormalizingFlow.sample_with_log_prob); only the chains' current states then go
                 through the density pass.  The reference evaluates log_prob on the proposals again
                 (monte_carlo.py:262) - the default here too.
        Returns the uint8 accept mask [B]."""
        B, dev = self.B, self.device
        if configs is None:
            z = self.nf_model.sample(B)
            configs = (z.reshape(B, self.num_particles, 2) + np.float32(self.half_width)).contiguous()
        configs = _lib.require_cuda(configs, "configs")
        if configs.dtype != torch.float32 or not configs.is_contiguous():
            raise _lib.FlowStateError("flowstate_b200: configurations must be contiguous float32")
        if logq is None and logq_new is not None:
            lq_old = self.nf_model.log_prob(self.centred(self.pos))
            lq_new = _lib.require_cuda(logq_new, "logq_new").float().contiguous()
        elif logq is None:
            both = torch.cat([self.centred(self.pos), self.centred(configs)], dim=0)
            lq = self.nf_model.log_prob(both)
            lq_old, lq_new = lq[:B].contiguous(), lq[B:].contiguous()
        else:
            lq_old, lq_new = logq
        mask = torch.empty(B, dtype=torch.uint8, device=dev)
        rng = self._rng_struct()
        if u is not None:
            u = _lib.require_cuda(u, "u")
        if self.fused_accept:
            # energy of the proposal + acceptance rule + masked update in one kernel (fs_accept_global_fused)
            E_new = torch.empty(B, dtype=torch.float32, device=dev)
            W_new = torch.empty(B, dtype=torch.float32, device=dev)
            Lx, Ly = self._L()
            _lib.check(_lib.lib().fs_accept_global_fused(
                _lib.ptr(self.pos), _lib.ptr(configs), _lib.ptr(self.E), _lib.ptr(self.W), _lib.ptr(E_new),
                _lib.ptr(W_new), _lib.ptr(lq_old), _lib.ptr(lq_new), _lib.ptr(u), C.byref(rng), float(self.beta),
                _lib.ptr(self.attempts), _lib.ptr(self.accepted), _lib.ptr(mask), B, self.num_particles,
                Lx, Ly, self._pot, _lib.stream_ptr(dev)))
            self.launches += 1
        else:
            E_new, W_new, _ = self.total_energy_virial(configs)
            self.accept_global(configs, E_new, W_new, lq_old, lq_new, u=u, mask=mask)
        self.proposal_energy = E_new
        return mask

    def accept_global(self, configs, E_new, W_new, lq_old, lq_new, u=None, mask=None):
        """Steps 3-5 of nf_big_move alone (fs_accept_global): acceptance rule on given proposal energies and flow
        log-densities, masked in-place update.  Returns the uint8 accept mask [B]."""
        B, dev = self.B, self.device
        if mask is None:
            mask = torch.empty(B, dtype=torch.uint8, device=dev)
        rng = self._rng_struct()
        _lib.check(_lib.lib().fs_accept_global(
            _lib.ptr(self.pos), _lib.ptr(configs), _lib.ptr(self.E), _lib.ptr(self.W), _lib.ptr(E_new),
            _lib.ptr(W_new), _lib.ptr(lq_old), _lib.ptr(lq_new), _lib.ptr(u), C.byref(rng), float(self.beta),
            _lib.ptr(self.attempts), _lib.ptr(self.accepted), _lib.ptr(mask), B, self.num_particles,
            _lib.stream_ptr(dev)))
        self.launches += 1
        return mask

    # -- observables ------------------------------------------------------
    def sample(self, cycle_number):
        """monte_carlo.py:416-444 vectorised: (cycle, E/N [B], rho, P [B], Lx, Ly, particles [B,N,2])."""
        volume = self.sim_box.volume
        density = self.num_particles / volume
        pressure = density / self.beta + self.W / (2.0 * volume)
        return (cycle_number, self.E / self.num_particles, density, pressure, self.sim_box.box_size_x,
                self.sim_box.box_size_y, self.pos.clone())

    @property
    def particles(self):
        return self.pos

    @property
    def attempts_displacement(self):
        return self.attempts

    @property
    def accepted_displacement(self):
        return self.accepted
