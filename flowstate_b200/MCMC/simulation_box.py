"""2-D periodic box.  Mirrors MCMC/simulation_box.py:3-65 of the reference
(same constructor, attributes and method names); the arithmetic runs in the
CUDA helpers fs_apply_pbc / fs_distances."""
import numpy as np
import torch

from ._bridge import _lib


def _dev():
    if not torch.cuda.is_available():
        raise _lib.FlowStateError("flowstate_b200: a CUDA device is required (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class SimulationBox:
    def __init__(self, box_size_x, box_size_y=None):
        if box_size_y is None:
            box_size_y = box_size_x
        self.box_size_x = box_size_x
        self.box_size_y = box_size_y
        self.volume = self.box_size_x * self.box_size_y

    # -- reference API ----------------------------------------------------
    def apply_pbc(self, position, checking=False):
        """position (2,) or (n, 2) -> wrapped copy (floor-mod per axis)."""
        arr = np.asarray(position)
        t = torch.as_tensor(arr.reshape(-1, 2), dtype=torch.float32).to(_dev()).contiguous()
        _lib.check(_lib.lib().fs_apply_pbc(_lib.ptr(t), t.shape[0], float(self.box_size_x),
                                           float(self.box_size_y), _lib.stream_ptr()))
        out = t.cpu().numpy().reshape(arr.shape)
        if checking:
            print("position % (box_size_x, box_size_y) =", out)
        return out

    def compute_distances(self, position_1, positions_2, checking=False):
        """Minimum-image distances from position_1 (2,) to positions_2 (n, 2)."""
        d = _dev()
        p1 = torch.as_tensor(np.asarray(position_1).reshape(2), dtype=torch.float32).to(d).contiguous()
        p2 = torch.as_tensor(np.asarray(positions_2).reshape(-1, 2), dtype=torch.float32).to(d).contiguous()
        r = torch.empty(p2.shape[0], dtype=torch.float32, device=d)
        _lib.check(_lib.lib().fs_distances(_lib.ptr(p1), 1, _lib.ptr(p2), p2.shape[0], float(self.box_size_x),
                                           float(self.box_size_y), _lib.ptr(r), _lib.stream_ptr()))
        return r.cpu().numpy().astype(np.float64)

    def compute_distance(self, position_1, position_2, checking=False):
        return float(self.compute_distances(position_1, np.asarray(position_2).reshape(1, 2))[0])

    def minimum_image(self, position_1, position_2, checking=False):
        """Displacement vector under the minimum-image convention (host convenience;
        only its norm is used on the sampling path)."""
        delta = np.asarray(position_1, dtype=np.float64) - np.asarray(position_2, dtype=np.float64)
        delta[0] -= self.box_size_x * np.round(delta[0] / self.box_size_x)
        delta[1] -= self.box_size_y * np.round(delta[1] / self.box_size_y)
        return delta
