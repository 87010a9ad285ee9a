"""flowstate_b200 - B200-native (sm_100a) NF-MCMC sampling hot path of flow-state.

Sub-packages mirror the reference's two import names so its drivers keep
working (`import normflows as NF; import MCMC as MC`,
hybrid_NF_MCMC/main_algorithm_1.py:29-30):

    import flowstate_b200.MCMC as MC
    import flowstate_b200.normflows as NF

All compute goes to hand-written CUDA kernels through the C ABI declared in
include/flowstate_b200.h; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import FlowStateError  # noqa: F401

__version__ = "0.1.0"
