"""Locates the ctypes binding whether this package is imported as `flowstate_b200.<name>` or,
drop-in style, as the top-level name the reference drivers use (`import MCMC` / `import normflows`)."""
try:
    from .. import _lib  # noqa: F401
except ImportError:      # imported as a top-level package: flowstate_b200 must be importable too
    from flowstate_b200 import _lib  # noqa: F401
