"""Import alias used ONLY by tests/reference_suite.py: lets the reference's own, unmodified test-suite
(`import torchsde`) exercise torchsde_b200 — the `sys.path` hook of the reference's tests (tests/test_sdeint.py:17).
Not part of the product."""
from torchsde_b200 import *  # noqa: F401,F403
from torchsde_b200 import __version__  # noqa: F401
