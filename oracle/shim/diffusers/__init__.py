"""TEST INFRASTRUCTURE — minimal restatement of the `diffusers==0.32.1` symbols that the
reference (tonyzyl/ladcast, pyproject.toml:34) imports on its ensemble-rollout hot path.

diffusers is a third-party dependency of the reference, absent from /root/reference and not
installable offline.  This package restates the *published* v0.32.1 behaviour of exactly the
pieces the reference calls (SURVEY.md Appendix A lists the call sites), so that the
UNMODIFIED reference files import and run on CPU when `oracle/shim` is first on sys.path.
PARITY UNPINNED for these pieces: there is no diffusers wheel offline to diff against; the
module *structure* is pinned by the exact parameter counts 374,938,452 / 1,605,496,660 /
256,411,145 (tests/test_oracle_shim.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from .schedulers.scheduling_edm_dpmsolver_multistep import EDMDPMSolverMultistepScheduler  # noqa: F401

__version__ = "0.32.1+oracle-shim"
