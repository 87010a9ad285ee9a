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
