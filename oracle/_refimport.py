"""Import harness for the UNMODIFIED reference (test infrastructure only).

Used only by oracle/make_golden.py in the build container, where /root/reference
exists.  Nothing on the GPU box imports this module (the reference tree does not
travel); the committed fixtures under tests/golden/ are what travels.

Recipe follows SURVEY.md section 8(c):
  * `normflows` imports as-is from <ref>/NF;
  * the MCMC modules import each other by bare name and need a module called
    `utils` exporting get_project_root() (MCMC/energy_calculator.py:4-6); the
    real MCMC/utils.py needs matplotlib/cycler, so a 3-function shim is put on
    sys.path first.
"""
import os
import sys
import tempfile

REF_ROOT = os.environ.get("FLOWSTATE_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "MCMC"))


_loaded = {}


def load():
    """Returns a dict of reference modules: normflows, simulation_box, potential,
    energy_calculator, monte_carlo."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    shim_dir = tempfile.mkdtemp(prefix="fs_ref_shim_")
    with open(os.path.join(shim_dir, "utils.py"), "w") as f:
        f.write(
            "def get_project_root():\n    return %r\n"
            "def set_icl_color_cycle(*a, **k):\n    pass\n"
            "def get_icl_heatmap_cmap(*a, **k):\n    return None\n" % REF_ROOT
        )
    sys.path.insert(0, shim_dir)
    sys.path.insert(1, os.path.join(REF_ROOT, "MCMC"))
    sys.path.insert(2, os.path.join(REF_ROOT, "NF"))
    import contextlib
    import io

    import normflows  # noqa
    import simulation_box  # noqa
    import potential  # noqa
    import energy_calculator  # noqa
    import monte_carlo  # noqa

    _loaded.update(
        normflows=normflows,
        simulation_box=simulation_box,
        potential=potential,
        energy_calculator=energy_calculator,
        monte_carlo=monte_carlo,
        quiet=lambda: contextlib.redirect_stdout(io.StringIO()),
    )
    return _loaded


def load_hybrid_utils():
    """The reference's hybrid_NF_MCMC/utils.py (analysis helpers).  It imports matplotlib / cycler for its
    plotting functions only; neither is installed here, so inert stand-ins are registered first."""
    if "hybrid_utils" in _loaded:
        return _loaded["hybrid_utils"]
    import importlib.util
    import types

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "cycler"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(cycler=lambda *a, **k: None, LinearSegmentedColormap=object, rcParams={})
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    spec = importlib.util.spec_from_file_location("fs_ref_hybrid_utils",
                                                  os.path.join(REF_ROOT, "hybrid_NF_MCMC", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded["hybrid_utils"] = mod
    return mod


class cuda_zeros_on_cpu:
    """The reference's SimpleLJ._energy creates its zero particle with device='cuda' unconditionally
    (NF/normflows/Energy/SimpleLJ.py:21) and moves it to x.device right after; on this GPU-less container the
    allocation itself fails.  Inside this context torch.zeros(..., device='cuda') allocates on the CPU instead -
    an environment stand-in (like the matplotlib stubs above), the reference's code runs unmodified."""

    def __enter__(self):
        import torch
        self._torch = torch
        self._orig = torch.zeros

        def zeros(*a, **k):
            if str(k.get("device", "")) == "cuda":
                k["device"] = "cpu"
            return self._orig(*a, **k)
        torch.zeros = zeros
        return self

    def __exit__(self, *exc):
        self._torch.zeros = self._orig
        return False
