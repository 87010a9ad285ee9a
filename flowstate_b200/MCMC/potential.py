"""Pair and external potentials.  Mirrors the live functions of
MCMC/potential.py (lennard_jones_energy_virial :3-29, double_well_potential
:55-116); evaluated by the CUDA helpers fs_lj_pair / fs_double_well."""
import numpy as np
import torch

from ._bridge import _lib
from .simulation_box import _dev


def lennard_jones_energy_virial(r, epsilon=1.0, sigma=1.0, cutoff_constant=2.5, shift=True):
    if epsilon != 1.0 or sigma != 1.0 or not shift:
        raise _lib.FlowStateError("flowstate_b200: only epsilon = sigma = 1 with the shifted cut-off is built "
                                  "(the only form the reference calls, energy_calculator.py:79-81)")
    arr = np.asarray(r)
    d = _dev()
    t = torch.as_tensor(arr.reshape(-1), dtype=torch.float32).to(d).contiguous()
    e = torch.empty_like(t)
    w = torch.empty_like(t)
    pot = _lib.make_pot(0, [0, 0], 1.0, 1.0, r_cut=cutoff_constant)
    _lib.check(_lib.lib().fs_lj_pair(_lib.ptr(t), t.numel(), pot, _lib.ptr(e), _lib.ptr(w), _lib.stream_ptr()))
    return (e.cpu().numpy().astype(np.float64).reshape(arr.shape),
            w.cpu().numpy().astype(np.float64).reshape(arr.shape))


def double_well_potential(position, box_size_x, box_size_y, V0_list=None, r0=1.0, k=10.0, num_wells=2):
    if V0_list is None:
        V0_list = [-4.0] * num_wells
    pos = np.atleast_2d(np.asarray(position))
    d = _dev()
    t = torch.as_tensor(pos, dtype=torch.float32).to(d).contiguous()
    v = torch.empty(t.shape[0], dtype=torch.float32, device=d)
    pot = _lib.make_pot(num_wells, V0_list, r0, k)
    _lib.check(_lib.lib().fs_double_well(_lib.ptr(t), t.shape[0], float(box_size_x), float(box_size_y), pot,
                                         _lib.ptr(v), _lib.stream_ptr()))
    out = v.cpu().numpy().astype(np.float64)
    return out[0] if out.shape[0] == 1 else out
