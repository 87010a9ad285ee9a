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


def lockstep_check(pos0, L, max_disp, pot, p_all, u_all, acc, idx=None, e=None, beta=1.0, tol_e=1e-5):
    """Replays one chain of a device run through the reference rule, step by step, on the SAME draws.

    pos0 (N,2) float32 start, p_all [S] particle indices, u_all [S,3] uniforms (two displacement uniforms and the accept
    uniform), acc [S] the device's decisions, idx / e optional device traces (particle index, (e_old, e_new)).
    The oracle state follows the device's decision after every step, so every decision is checked against
    MCMC/monte_carlo.py:191-223 on the very state the device saw.

    Tolerances.  Energies: |e_dev - e_ref| <= tol_e * max(1, M) with M = sum_j |LJ(r_pj)| + |V_ext| the magnitude of the
    summed terms (= |e| when the terms share a sign; a float32 sum of cancelling pair terms cannot be more accurate than
    that, and neither is the reference itself once its state is float32).  A decision may differ from the
    rule only inside the epsilon band |log u - Delta| <= beta tol_e (M_old + M_new) (SURVEY.md 7.2 with |e| -> M), or
    when a pair sits within 1e-6 of the hard-core radius (the device compares r^2, the reference r).
    Returns dict(flips_in_band, outside_band, max_energy_err, final (N,2) float32 state)."""
    state = np.array(pos0, dtype=np.float32)
    flips = outside = 0
    max_err = 0.0
    for s in range(len(acc)):
        p = int(p_all[s])
        if idx is not None and int(idx[s]) != p:
            outside += 1
        truth = state.astype(np.float64)
        eno, _ = er.particle_energy_virial(truth, p, L, L, pot)
        new = state.copy()
        new[p] += (np.asarray(u_all[s, :2], dtype=np.float64) - 0.5) * max_disp    # monte_carlo.py:161-163
        new[p] = er.apply_pbc(new[p], L, L)
        new64 = new.astype(np.float64)
        enn, _ = er.particle_energy_virial(new64, p, L, L, pot)
        mo = er.particle_energy_magnitude(truth, p, L, L, pot)
        mn = er.particle_energy_magnitude(new64, p, L, L, pot)
        core_band = False
        if len(state) > 1:
            dmin_o = np.min(er.distances(truth[p], np.delete(truth, p, 0), L, L))
            dmin_n = np.min(er.distances(new64[p], np.delete(new64, p, 0), L, L))
            core_band = abs(dmin_o - er.R_CORE) < 1e-6 or abs(dmin_n - er.R_CORE) < 1e-6
        if e is not None and not core_band:
            for got, ref, mag in ((float(e[s][0]), eno, mo), (float(e[s][1]), enn, mn)):
                if np.isinf(got) != np.isinf(ref):
                    outside += 1
                elif np.isfinite(ref):
                    max_err = max(max_err, abs(got - ref) / max(1.0, mag))
        eps = tol_e * ((mo if np.isfinite(mo) else 0.0) + (mn if np.isfinite(mn) else 0.0))
        if enn <= eno:
            ref_ok = True
            band = bool(np.isfinite(eno)) and abs(enn - eno) <= eps
        elif np.isinf(enn):
            ref_ok, band = False, False
        else:
            delta = -beta * (enn - eno)
            u = float(u_all[s, 2])
            ref_ok = bool(u < np.exp(delta))
            band = (u > 0 and abs(np.log(u) - delta) <= beta * eps) or abs(enn - eno) <= eps
        if bool(acc[s]) != ref_ok:
            if band or core_band:
                flips += 1
            else:
                outside += 1
        if acc[s]:
            state = new
    return {"flips_in_band": flips, "outside_band": outside, "max_energy_err": max_err, "final": state}
