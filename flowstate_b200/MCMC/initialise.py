"""Initial configurations of the chains (host-side setup).

`initialise_low_left` / `initialise_low_right` / `initialise_fcc` restate the reference's lattice builders
(MCMC/initialise.py:118-305, 8-116) without their matplotlib parts: same arguments, same geometry, same particle
order, same return value `(particles (N, 2) float64, SimulationBox)`.  `initialise_chains` builds the start of B chains
at once the way the drivers do (even chains in the left well, odd chains in the right one,
hybrid_NF_MCMC/main_algorithm_1.py:149-165).  For more than 12 particles the reference's low initialisers raise; the
drivers then use `initialise_cluster` (the same grid rule without the limit) and the benchmarks the seeded jittered
lattice of SURVEY.md 8d.
"""
import math

import numpy as np

from .simulation_box import SimulationBox


def _box(num_particles, rho, aspect_ratio=1.0):
    area = num_particles / rho
    return SimulationBox(float(np.sqrt(area * aspect_ratio)), float(np.sqrt(area / aspect_ratio)))


def _low_grid(num_particles, centre, box):
    """MCMC/initialise.py:163-205: grid_cols = ceil(sqrt N), grid_rows = ceil(N / cols), spacing min(1.5, box_x /
    (2 (cols - 1)), box_y / (rows - 1)), rows along y, filled row by row, centred on `centre`, wrapped into the box."""
    if num_particles == 1:
        return np.array([[centre[0], centre[1]]], dtype=np.float64)
    cols = int(math.ceil(math.sqrt(num_particles)))
    rows = int(math.ceil(num_particles / cols))
    max_sep_x = box.box_size_x / (2 * (cols - 1)) if cols > 1 else float("inf")
    max_sep_y = box.box_size_y / (rows - 1) if rows > 1 else float("inf")
    spacing = min(1.5, max_sep_x, max_sep_y)
    width, height = (cols - 1) * spacing, (rows - 1) * spacing
    k = np.arange(num_particles)
    x = centre[0] - width / 2 + (k % cols) * spacing
    y = centre[1] - height / 2 + (k // cols) * spacing
    return np.stack([x % box.box_size_x, y % box.box_size_y], axis=1)        # SimulationBox.apply_pbc (:19-29)


def _check_low(num_particles):
    if num_particles < 1 or num_particles > 12:
        raise ValueError("Number of particles for low initialization must be between 1 and 12.")


def initialise_low_left(num_particles=2, rho=0.5, aspect_ratio=1.0, visualise=False, checking=False):
    """MCMC/initialise.py:118-211: 1..12 particles on a grid centred on the left well (Lx / 4, Ly / 2)."""
    _check_low(num_particles)
    box = _box(num_particles, rho, aspect_ratio)
    return _low_grid(num_particles, (box.box_size_x / 4, box.box_size_y / 2), box), box


def initialise_low_right(num_particles=2, rho=0.5, aspect_ratio=1.0, visualise=False, checking=False):
    """MCMC/initialise.py:214-305: the same around the right well (3 Lx / 4, Ly / 2)."""
    _check_low(num_particles)
    box = _box(num_particles, rho, aspect_ratio)
    return _low_grid(num_particles, (3 * box.box_size_x / 4, box.box_size_y / 2), box), box


def initialise_cluster(num_particles, rho, side="left", aspect_ratio=1.0):
    """The low initialisers' grid rule for any particle count (the reference stops at 12)."""
    box = _box(num_particles, rho, aspect_ratio)
    cx = box.box_size_x / 4 if side == "left" else 3 * box.box_size_x / 4
    return _low_grid(num_particles, (cx, box.box_size_y / 2), box), box


def initialise_fcc(num_particles=48, rho=0.5, aspect_ratio=1.5, visualise=False, checking=False):
    """MCMC/initialise.py:8-116: two interpenetrating rectangular sublattices ((i dx, j dy) and ((i + .5) dx,
    (j + .5) dy), nx = ceil(sqrt(N / 2 * aspect)), ny = ceil(N / (2 nx)), dx = Lx / (nx - .5), dy = Ly / (ny - .5));
    the N candidates closest to the box centre are kept (stable order of np.argsort on the squared distances)."""
    box = _box(num_particles, rho, aspect_ratio)
    nx = math.ceil(np.sqrt(num_particles / 2 * aspect_ratio))
    ny = math.ceil(num_particles / (2 * nx))
    dx = box.box_size_x / (nx - 0.5)
    dy = box.box_size_y / (ny - 0.5)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    a = np.stack([i * dx, j * dy], axis=-1).reshape(-1, 2)
    b = np.stack([(i + 0.5) * dx, (j + 0.5) * dy], axis=-1).reshape(-1, 2)
    cand = np.empty((2 * nx * ny, 2), dtype=np.float64)
    cand[0::2], cand[1::2] = a, b                                            # A and B of a cell are appended in turn
    cand[:, 0] %= box.box_size_x
    cand[:, 1] %= box.box_size_y
    centre = np.array([box.box_size_x / 2, box.box_size_y / 2])
    order = np.argsort(np.sum((cand - centre) ** 2, axis=1))
    return cand[order[:num_particles]], box


def initialise_chains(chains, num_particles, rho, first_chain=0, aspect_ratio=1.0, dtype=np.float32):
    """Start configurations of `chains` chains with global ids first_chain ..: even ids in the left well, odd ids in
    the right one (main_algorithm_1.py:149-165).  Returns ((chains, N, 2) array, SimulationBox)."""
    left, box = initialise_cluster(num_particles, rho, "left", aspect_ratio)
    right, _ = initialise_cluster(num_particles, rho, "right", aspect_ratio)
    ids = first_chain + np.arange(chains)
    out = np.where((ids % 2 == 0)[:, None, None], left[None], right[None])
    return out.astype(dtype), box


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
