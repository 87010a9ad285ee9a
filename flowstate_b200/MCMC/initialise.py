"""Host-side initial configurations (setup, not on the hot path).

The reference's lattice initialisers (MCMC/initialise.py:118-305) are plotting-
entangled host code and out of scope as compute;
these are small synthetic stand-ins with the same call shape and return value
`(particles (N,2) float64, SimulationBox)`: a compact grid of spacing 1.5 placed
on the left / right well centre, and the seeded jittered lattice used by the
benchmarks."""
import numpy as np

from .simulation_box import SimulationBox


def _box(num_particles, rho, aspect_ratio=1.0):
    area = num_particles / rho
    lx = float(np.sqrt(area * aspect_ratio))
    ly = float(area / lx)
    return SimulationBox(lx, ly)


def _grid(num_particles, centre, spacing=1.5):
    m = int(np.ceil(np.sqrt(num_particles)))
    g = np.array([(i, j) for i in range(m) for j in range(m)][:num_particles], dtype=np.float64)
    return (g - g.mean(axis=0)) * spacing + np.asarray(centre)


def initialise_low_left(num_particles, rho, aspect_ratio=1.0, visualise=False, checking=False):
    box = _box(num_particles, rho, aspect_ratio)
    return _grid(num_particles, (box.box_size_x / 4, box.box_size_y / 2)), box


def initialise_low_right(num_particles, rho, aspect_ratio=1.0, visualise=False, checking=False):
    box = _box(num_particles, rho, aspect_ratio)
    return _grid(num_particles, (3 * box.box_size_x / 4, box.box_size_y / 2)), box


def jittered_lattice(num_particles, rho, seed, jitter=0.1, batch=None):
    """Seeded jittered square lattice, float32-exact.  batch=None -> (N,2); else (batch,N,2)
    with seeds seed, seed+1, ...  Returns (positions float32, box length)."""
    L = float(np.float32(np.sqrt(num_particles / rho)))
    m = int(np.ceil(np.sqrt(num_particles)))
    a = L / m
    sites = np.array([((i + 0.5) * a, (j + 0.5) * a) for i in range(m) for j in range(m)][:num_particles])

    def one(s):
        return (sites + (np.random.default_rng(s).random((num_particles, 2)) - 0.5) * jitter * a).astype(np.float32)
    if batch is None:
        return one(seed), L
    return np.stack([one(seed + b) for b in range(batch)]), L
