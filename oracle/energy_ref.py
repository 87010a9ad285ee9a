"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the reference pair energy.

Follows, function by function (paths relative to the reference root):
  MCMC/simulation_box.py:19-65     apply_pbc / minimum_image / compute_distances
  MCMC/potential.py:3-29           lennard_jones_energy_virial (cut 2.5, shifted)
  MCMC/potential.py:55-116         double_well_potential (tanh wells)
  MCMC/energy_calculator.py:48-108 calculate_particle_energy_virial
  MCMC/energy_calculator.py:121-203 calculate_total_energy_virial

The reference walks pairs in a Python loop; here the same arithmetic is applied
to whole rows with numpy, keeping the reference's dtype flow: pair geometry
(difference, minimum image, norm) runs in the dtype of the position array
(float64 normally, float32 after an NF acceptance - SURVEY.md 7.2), the LJ
formula and the well term run in float64.

Pinned against reference outputs in tests/golden/energy_*.npz.
"""
import numpy as np

R_CORE = 0.5   # energy_calculator.py:73,150  (r < 0.5 -> inf)
R_CUT = 2.5    # potential.py:3  cutoff_constant


class Potential:
    """Parameters the reference passes around as loose arguments
    (energy_calculator.py:10-21)."""

    def __init__(self, num_wells=2, V0_list=(-10.0, -10.5), r0=1.2, k=15.0):
        self.num_wells = int(num_wells)
        self.V0_list = list(V0_list)
        self.r0 = float(r0)
        self.k = float(k)


def apply_pbc(position, Lx, Ly):
    """simulation_box.py:19-29: Python/numpy floor-mod per axis."""
    return np.array([position[0] % Lx, position[1] % Ly])


def distances(p1, p2s, Lx, Ly):
    """simulation_box.py:31-65: minimum-image distances from p1 to rows of p2s.
    Geometry in the dtype of the inputs, result stored as float64 (the reference
    writes each distance into an np.zeros float64 buffer, :62-64)."""
    delta = p1[None, :] - p2s
    dx = delta[:, 0]
    dy = delta[:, 1]
    dx = dx - Lx * np.round(dx / Lx)          # round-half-even, :38
    dy = dy - Ly * np.round(dy / Ly)          # :39
    r = np.sqrt(dx * dx + dy * dy)            # np.linalg.norm, :53
    return r.astype(np.float64)


def lj_energy_virial(r):
    """potential.py:3-29 with epsilon=sigma=1, cutoff 2.5, shift=True."""
    r = np.asarray(r, dtype=np.float64)
    energy = np.zeros_like(r)
    virial = np.zeros_like(r)
    mask = r <= R_CUT
    sr6 = (1.0 / r[mask]) ** 6
    sr12 = sr6 * sr6
    energy[mask] = 4.0 * (sr12 - sr6)
    virial[mask] = 48.0 * (sr12 - 0.5 * sr6)
    sr6_cut = (1.0 / R_CUT) ** 6
    energy_cut = 4.0 * (sr6_cut * sr6_cut - sr6_cut)
    energy[mask] -= energy_cut
    return energy, virial


def double_well(position, Lx, Ly, pot):
    """potential.py:55-116.  Returns an array (N,) for (N,2) input, scalar for (2,)."""
    position = np.atleast_2d(position)
    x = position[:, 0]
    y = position[:, 1]
    centers = []
    if pot.num_wells >= 1:
        centers.append([Lx / 4, Ly / 2])
    if pot.num_wells == 2:
        centers.append([3 * Lx / 4, Ly / 2])
    centers = np.array(centers, dtype=np.float64).reshape(-1, 2)
    V = np.zeros(x.shape, dtype=np.float64)
    for i, c in enumerate(centers):
        dx = x - c[0]                       # float64 (c[0] is np.float64)
        dy = y - c[1]
        dx = dx - Lx * np.round(dx / Lx)
        dy = dy - Ly * np.round(dy / Ly)
        r = np.sqrt(dx ** 2 + dy ** 2)
        transition = 0.5 * (1 + np.tanh(pot.k * (r - pot.r0)))
        V += pot.V0_list[i] * (1 - transition)
    if V.shape[0] == 1:
        return V[0]
    return V


def particle_energy_virial(positions, p, Lx, Ly, pot):
    """energy_calculator.py:48-108: sum_{j != p} LJ(r_pj) + V_ext(p); any
    r < 0.5 -> (inf, inf).  Virial carries no external term."""
    others = np.delete(positions, p, axis=0)
    r = distances(positions[p], others, Lx, Ly)
    if np.any(r < R_CORE):
        return float("inf"), float("inf")
    e, w = lj_energy_virial(r)
    pe = np.sum(e)
    pw = np.sum(w)
    if pot.num_wells > 0:
        pe = pe + double_well(positions[p], Lx, Ly, pot)
    return float(pe), float(pw)


def particle_energy_magnitude(positions, p, Lx, Ly, pot):
    """sum_j |LJ(r_pj)| + |V_ext(p)|: the magnitude of the terms particle_energy_virial adds up.  A float32 evaluation
    of the pair terms is accurate relative to THIS quantity (the total can be much smaller when attractive and
    repulsive terms cancel); the parity tests scale their 1e-5 tolerance and the acceptance epsilon band with it."""
    others = np.delete(positions, p, axis=0)
    r = distances(positions[p], others, Lx, Ly)
    if np.any(r < R_CORE):
        return float("inf")
    e, _ = lj_energy_virial(r)
    m = np.sum(np.abs(e))
    if pot.num_wells > 0:
        m = m + abs(double_well(positions[p], Lx, Ly, pot))
    return float(m)


def total_energy_virial(positions, Lx, Ly, pot):
    """energy_calculator.py:121-203: sum_{i<j} LJ + sum_i V_ext; the reference
    returns (inf, inf) at the first row holding a pair closer than 0.5."""
    n = len(positions)
    E = 0.0
    W = 0.0
    for i in range(n - 1):
        r = distances(positions[i], positions[i + 1:], Lx, Ly)
        if np.any(r < R_CORE):
            return float("inf"), float("inf")
        e, w = lj_energy_virial(r)
        E += np.sum(e)
        W += np.sum(w)
    if pot.num_wells > 0:
        E += np.sum(double_well(positions, Lx, Ly, pot))
    return float(E), float(W)


# ---------------------------------------------------------------------------
# Synthetic inputs: seeded jittered lattices, float32-exact.
# ---------------------------------------------------------------------------

def box_length(n, rho):
    """L = sqrt(N/rho) rounded to float32 so both sides see the same box."""
    return float(np.float32(np.sqrt(n / rho)))


def jittered_lattice(n, rho, seed, jitter=0.1, dtype=np.float32):
    """SURVEY.md Appendix B recipe: m=ceil(sqrt N), a=L/m, sites (i+.5,j+.5)a,
    jitter U(-jitter/2, jitter/2)*a.  Values are float32-representable."""
    L = box_length(n, rho)
    m = int(np.ceil(np.sqrt(n)))
    a = L / m
    pts = np.array([((i + 0.5) * a, (j + 0.5) * a) for i in range(m) for j in range(m)][:n])
    pts = pts + (np.random.default_rng(seed).random((n, 2)) - 0.5) * jitter * a
    pts = pts.astype(np.float32)
    return pts.astype(dtype), L


def batch_lattices(B, n, rho, seed0, jitter=0.1):
    out = np.empty((B, n, 2), dtype=np.float32)
    L = box_length(n, rho)
    for b in range(B):
        out[b], _ = jittered_lattice(n, rho, seed0 + b, jitter)
    return out, L
