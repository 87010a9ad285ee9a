"""Drop-in for the reference's vendored `normflows` (hot-path subset).

Mirrors the module interface used by the drivers
(hybrid_NF_MCMC/main_algorithm_1.py:277-284): NF.NormalizingFlow,
NF.flows.CircularCoupledRationalQuadraticSpline, NF.Energy.UniformParticle, with
the reference's parameter/buffer names so its .pth files load
unchanged.  In eval mode on CUDA tensors every forward / inverse / log_prob /
sample runs in the CUDA kernels (fs_flow_inverse / fs_flow_forward); train mode
is the autograd path used for (Algorithm 2) training.
"""
from .core import NormalizingFlow
from . import core, flows, Energy, nets, utils

__version__ = "1.7.3+b200"
__all__ = ["NormalizingFlow", "core", "flows", "Energy", "nets", "utils"]
