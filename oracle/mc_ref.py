"""Oracle (TEST INFRASTRUCTURE): restatement of the reference Metropolis sampler.

Follows MCMC/monte_carlo.py of the reference:
  :146-189  particle_displacement        -> ChainRef.local_step
  :191-223  metropolis_acceptance_particle_move
  :235-303  nf_big_move                  -> ChainRef.global_move
  :375-403  adjust_displacement
  :416-444  sample

RNG draw order per local step: integers(N), random(2), and a
third random() ONLY on finite uphill moves.  The generator is injected so tests
can use numpy's PCG64 (what the reference uses, monte_carlo.py:92-95), a spy
that records draws, or a replayer.

Pinned against reference traces in tests/golden/mc_*.npz.
"""
import numpy as np

from . import energy_ref as er


class SpyRNG:
    """Wraps a numpy Generator and records every draw (for replay tests)."""

    def __init__(self, rng):
        self.rng = rng
        self.ints = []
        self.uniforms = []

    def integers(self, n):
        v = self.rng.integers(n)
        self.ints.append(int(v))
        return v

    def random(self, k=None):
        v = self.rng.random(k)
        if k is None:
            self.uniforms.append(float(v))
        else:
            self.uniforms.extend(float(x) for x in v)
        return v


class ReplayRNG:
    """Replays recorded draws: one cursor over particle indices, one over uniforms."""

    def __init__(self, ints, uniforms):
        self.ints = list(ints)
        self.uniforms = list(uniforms)
        self.ci = 0
        self.cu = 0

    def integers(self, n):
        v = self.ints[self.ci]
        self.ci += 1
        return v

    def random(self, k=None):
        if k is None:
            v = self.uniforms[self.cu]
            self.cu += 1
            return v
        v = np.array(self.uniforms[self.cu:self.cu + k], dtype=np.float64)
        self.cu += k
        return v


class ChainRef:
    """One Markov chain with the reference's state and counters
    (monte_carlo.py:64-126)."""

    def __init__(self, particles, L, temperature, pot, max_displacement=0.65,
                 target_acceptance=0.5, rng=None):
        self.particles = particles
        self.Lx = self.Ly = L
        self.half_width = L / 2
        self.beta = 1.0 / temperature
        self.n = len(particles)
        self.pot = pot
        self.max_displacement = max_displacement
        self.target_acceptance = target_acceptance
        self.attempts = 0
        self.accepted = 0
        self.prev_attempts = 0
        self.prev_accepted = 0
        self.rng = rng if rng is not None else np.random.default_rng()
        self.E, self.W = er.total_energy_virial(particles, L, L, pot)

    # -- local move -------------------------------------------------------
    def metropolis(self, old_e, new_e):
        """monte_carlo.py:191-223.  Returns (accept, u or None)."""
        if new_e <= old_e:
            return True, None
        if np.isinf(new_e):
            return False, None
        factor = np.exp(-self.beta * (new_e - old_e))
        u = self.rng.random()
        return bool(u < factor), float(u)

    def local_step(self):
        """monte_carlo.py:146-179.  Returns a trace tuple
        (p, e_old, e_new, accepted, u)."""
        self.attempts += 1
        p = int(self.rng.integers(self.n))
        eno, viro = er.particle_energy_virial(self.particles, p, self.Lx, self.Ly, self.pot)
        disp = (self.rng.random(2) - 0.5) * self.max_displacement
        new = self.particles.copy()
        new[p] += disp
        new[p] = er.apply_pbc(new[p], self.Lx, self.Ly)
        enn, virn = er.particle_energy_virial(new, p, self.Lx, self.Ly, self.pot)
        acc, u = self.metropolis(eno, enn)
        if acc:
            self.particles = new
            self.accepted += 1
            self.E += enn - eno
            self.W += virn - viro
        return p, eno, enn, acc, u

    # -- global move ------------------------------------------------------
    def global_move(self, config, nll_old, nll_new):
        """monte_carlo.py:235-303 with the two flow negative log-likelihoods
        supplied by the caller (the reference gets them from nf_model.log_prob
        at :261-262).  Returns (accept, ratio_log, u)."""
        self.attempts += 1
        eno = self.E
        enn, virn = er.total_energy_virial(config, self.Lx, self.Ly, self.pot)
        ratio_log = -self.beta * (enn - eno) - (nll_new - nll_old)
        with np.errstate(over="ignore", invalid="ignore"):
            ratio = np.exp(ratio_log)
        u = None
        if ratio >= 1.0:
            accept = True
        else:
            u = float(self.rng.random())
            accept = bool(u < ratio)
        if accept:
            self.particles = config
            self.accepted += 1
            self.E, self.W = enn, virn
        else:
            # :299-301 the reference recomputes the old total to restore its cache
            self.E, self.W = er.total_energy_virial(self.particles, self.Lx, self.Ly, self.pot)
        return accept, float(ratio_log), u

    # -- adaptation / observables ----------------------------------------
    def adjust_displacement(self):
        """monte_carlo.py:375-403."""
        if self.attempts > self.prev_attempts:
            d_att = self.attempts - self.prev_attempts
            d_acc = self.accepted - self.prev_accepted
            frac = d_acc / d_att if d_att > 0 else 0
            new = self.max_displacement * (frac / self.target_acceptance)
            ratio = new / self.max_displacement
            if ratio > 1.5:
                new = self.max_displacement * 1.5
            elif ratio < 0.5:
                new = self.max_displacement * 0.5
            self.max_displacement = new
            self.prev_attempts = self.attempts
            self.prev_accepted = self.accepted

    def sample(self, cycle):
        """monte_carlo.py:416-444 (without the history appends)."""
        volume = self.Lx * self.Ly
        density = self.n / volume
        pressure = density / self.beta + self.W / (2.0 * volume)
        return (cycle, self.E / self.n, density, pressure, self.Lx, self.Ly,
                self.particles.copy())
