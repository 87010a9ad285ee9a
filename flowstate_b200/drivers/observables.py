"""On-device observables of sampled configurations, mirroring the reference's analysis helpers.

Same names, arguments and return values as hybrid_NF_MCMC/utils.py:
  classify_particles(positions, halfbox, r0)                      (:107-141)
  calculate_well_statistics(configurations, start_idx, half_box, r0=1.2)   (:61-104)
  calculate_pair_correlation(final_samples, n_particles, bound, dr=None)   (:530-556)
plus the per-run output files of the drivers (main_algorithm_1.py:499-548).
The O(B N) classification and the O(B N^2) distance histogram run in fs_classify_wells /
fs_pair_histogram; the cumulative statistics and the normalisation are the reference's float64
host arithmetic on B-sized arrays.
"""
import csv
import os

import numpy as np
import torch

from .. import _lib


def _device_configs(configurations):
    t = configurations if torch.is_tensor(configurations) else torch.as_tensor(np.asarray(configurations))
    if t.dim() == 2:
        t = t[None]
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.FlowStateError("flowstate_b200: observables need a CUDA device (no CPU fallback)")
        t = t.cuda()
    return t.to(torch.float32).contiguous()


def _classify(configurations, half_box, r0):
    pos = _device_configs(configurations)
    B, N = pos.shape[0], pos.shape[1]
    dev = pos.device
    cls = torch.empty(B, N, dtype=torch.uint8, device=dev)
    state = torch.empty(B, dtype=torch.uint8, device=dev)
    avg_x = torch.empty(B, dtype=torch.float64, device=dev)
    L = float(half_box) * 2
    _lib.check(_lib.lib().fs_classify_wells(_lib.ptr(pos), B, N, L, L, float(r0), _lib.ptr(cls), _lib.ptr(state),
                                            _lib.ptr(avg_x), _lib.stream_ptr(dev)))
    return cls, state, avg_x


def classify_particles(positions, halfbox, r0):
    """(n_configs, N) array of 'A' / 'B' / 'Outside' (utils.py:107-141)."""
    cls, _, _ = _classify(positions, halfbox, r0)
    names = np.array(["Outside", "A", "B"])
    return names[cls.cpu().numpy()]


def calculate_well_statistics(configurations, start_idx, half_box, r0=1.2):
    """avg_x, p_a, p_b, deltaF, runs as lists, cumulative over configurations[start_idx:] (utils.py:61-104)."""
    _, state, avg_x = _classify(configurations, half_box, r0)
    state = state.cpu().numpy()[start_idx:]
    avg_x = avg_x.cpu().numpy()[start_idx:]
    n = len(state)
    runs = np.arange(1, n + 1)
    p_a = np.cumsum(state == 1) / runs
    p_b = np.cumsum(state == 2) / runs
    both = (p_a > 0) & (p_b > 0)
    delta_f = np.zeros(n)
    delta_f[both] = np.log(p_b[both] / p_a[both])
    return list(avg_x), list(p_a), list(p_b), list(delta_f), list(runs)


def pair_histogram(final_samples, bound, dr):
    """Per-configuration np.histogram(distances, np.arange(0, bound + dr, dr)) of the minimum-image pair
    distances (both orders counted, zeros dropped): uint32 tensor [n_configs, nbins] on the device."""
    cfg = _device_configs(final_samples)
    B, N = cfg.shape[0], cfg.shape[1]
    edges = np.arange(0, bound + dr, dr)
    nbins = len(edges) - 1
    counts = torch.empty(B, nbins, dtype=torch.int32, device=cfg.device)
    _lib.check(_lib.lib().fs_pair_histogram(_lib.ptr(cfg), B, N, float(bound), float(dr), nbins, _lib.ptr(counts),
                                            _lib.stream_ptr(cfg.device)))
    return counts


def calculate_pair_correlation(final_samples, n_particles, bound, dr=None):
    """(r values, g(r)) averaged over the samples (utils.py:530-556)."""
    if dr is None:
        dr = bound / 50
    counts = pair_histogram(final_samples, bound, dr).cpu().numpy().astype(np.int64)
    norm = n_particles * (n_particles - 1) / 2
    rou = n_particles / (4 * bound * bound)
    i_vals = np.arange(0, bound, dr)
    area = np.pi * ((i_vals + dr) ** 2 - i_vals ** 2)
    result = counts / (norm * rou * area)
    g_r = result.mean(axis=0)
    try:
        import pandas as pd
        g_r = pd.Series(g_r)
    except ImportError:
        pass
    return i_vals, g_r


def save_run_outputs(run_folder, local_samples, testing_samples=None):
    """sampled_data.csv + mc_run_configs.npy (+ mc_run_testing_configs.npy) of one chain, in the reference's
    formats (main_algorithm_1.py:499-548).  local_samples: list of MonteCarlo.sample() tuples."""
    os.makedirs(run_folder, exist_ok=True)
    with open(os.path.join(run_folder, "sampled_data.csv"), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["cycle_number", "energy_per_particle", "density", "pressure", "box_size_x", "box_size_y",
                     "particle_configuration"])
        for (cycle, epp, rho, pres, lx, ly, particles) in local_samples:
            wr.writerow([cycle, epp, rho, pres, lx, ly, np.array(particles).flatten().tolist()])
    np.save(os.path.join(run_folder, "mc_run_configs.npy"), np.array([np.array(s[6]) for s in local_samples]))
    if testing_samples is not None:
        np.save(os.path.join(run_folder, "mc_run_testing_configs.npy"), np.array(testing_samples))
