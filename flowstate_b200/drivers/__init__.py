"""Batched GPU versions of the reference's experiment drivers (hybrid_NF_MCMC/main_algorithm_{1,2}.py)."""
from .hybrid import run_algorithm_1, run_mcmc_only, run_algorithm_2, HybridConfig  # noqa: F401
